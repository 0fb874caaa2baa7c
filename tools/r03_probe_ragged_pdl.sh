#!/bin/bash
# ragged_pdl experiment (NOT KEPT: the switch is no longer in the code): parity tests with the option on, then configs[3] A/B on ONE box
# (B200TTS_RAGGED_PDL = 0 / 1). The change was `PdlScope pdl_short(m.U == 1 || (m.ragged && e.ragged_pdl))` around the per-utterance conv-pos
# launches of f5_steps and `PdlScope(e.ragged_pdl)` around the per-utterance Euler launches. Result: profiles/r02/probe_zp_ragged_pdl_experiment.log.
OUT=gpurun_out/r03_probe_ragged_pdl.log
: > $OUT
echo "== pytest ragged tests, B200TTS_RAGGED_PDL=1" >> $OUT
B200TTS_RAGGED_PDL=1 timeout 100 python -m pytest tests/test_gpu_f5.py -q -x -k "ragged" 2>&1 | tail -4 >> $OUT
for flag in 0 1; do
  echo "== bench config4, B200TTS_RAGGED_PDL=$flag" >> $OUT
  B200TTS_RAGGED_PDL=$flag timeout 80 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>>$OUT | python -c "
import json,sys; d=json.loads(sys.stdin.read()); pm=d['profile_ms']
print('ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), 'MHz', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'],
      {k: round(pm[k],1) for k in ('f5.embed_x','f5.cast','f5.conv_pos','f5.euler','f5.chain','f5.attention')})" >> $OUT
done
cat $OUT

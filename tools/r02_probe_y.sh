#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-y}.log
: > $OUT
echo "== ragged == singles test" >> $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_f5.py tests/test_gpu_fullsize.py -m gpu -q -k "ragged or pipeline" 2>&1 | tail -5 >> $OUT
for rows in 0 -1 9472 37888; do
  echo "== config4 microbatch rows=$rows" >> $OUT
  if [ "$rows" = "-1" ]; then unset B200TTS_MICROBATCH_ROWS; else export B200TTS_MICROBATCH_ROWS=$rows; fi
  timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_y.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ['value','ms_per_step','clocks']}), json.dumps(d['e2e']['ms_per_step']), d['roofline']['frac'])
" >> $OUT 2>&1
done
cat $OUT

#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
for cfg in "2 3" "4 2" "2 2"; do
  set -- $cfg
  echo "== config4 ATTN_G=$1 ATTN_POLY=$2" >> $OUT
  B200TTS_ATTN_G=$1 B200TTS_ATTN_POLY=$2 timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
pm=d['profile_ms']
print(json.dumps({'ms_per_step':d['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'attn':round(pm['f5.attention']),'chain':round(pm['f5.chain']),'clk':d['clocks']['sm_mhz']}))
" >> $OUT 2>&1
done
echo "== attention tests POLY=3" >> $OUT
B200TTS_ATTN_POLY=3 timeout -s KILL 300 python -m pytest tests/test_gpu_f5.py -m gpu -q -k attention 2>&1 | tail -2 >> $OUT
cat $OUT

#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
for p in 1 0 2 1; do
  echo "== config4 ATTN_POLY=$p" >> $OUT
  B200TTS_ATTN_POLY=$p timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
pm=d['profile_ms']
print(json.dumps({'ms_per_step':d['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'attn':round(pm['f5.attention']),'chain':round(pm['f5.chain']),'clk':d['clocks']['sm_mhz']}))
" >> $OUT 2>&1
done
cat $OUT

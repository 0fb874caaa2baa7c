#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
echo "== pytest f5 + fullsize" >> $OUT
timeout -s KILL 1200 python -m pytest tests/test_gpu_f5.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -6 >> $OUT
for u in 8 9; do
  echo "== pipeline (uniform) utterances=$u" >> $OUT
  timeout -s KILL 600 python bench.py --workload pipeline --utterances $u --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
u=$u
pm=d['profile_ms']
print(json.dumps({'ms_per_step':d['ms_per_step'],'ms_per_utt':d['ms_per_step']/u,'chain_per_utt':pm['f5.chain']/u,'chain_launch_ms':d['roofline']['avg_launch_ms'],'frac':d['roofline']['frac']}))
" >> $OUT 2>&1
done
for u in 8 16 32 64; do
  echo "== config4 (ragged) utterances=$u" >> $OUT
  timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --config4-utterances $u 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
u=$u
pm=d['profile_ms']
print(json.dumps({'ms_per_step':d['ms_per_step'],'ms_per_utt':d['ms_per_step']/u,'chain_per_utt':pm['f5.chain']/u,'attn_per_utt':pm['f5.attention']/u,'chain_launch_ms':d['roofline']['avg_launch_ms'],'frac':d['roofline']['frac'],'clk':d['clocks']['sm_mhz']}))
" >> $OUT 2>&1
done
cat $OUT

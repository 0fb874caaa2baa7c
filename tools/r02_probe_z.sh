#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
echo "== bvg_branches" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py bvg_branches 2>&1 | grep -v done >> $OUT
echo "== bench f5" >> $OUT
timeout -s KILL 600 python bench.py --workload f5 --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
pm=d['profile_ms']
print(json.dumps({'ms_per_step':d['ms_per_step'], **{k:round(v,2) for k,v in pm.items() if v>0.5}}))
" >> $OUT 2>&1
echo "== pytest bigvgan + f5 + fullsize" >> $OUT
timeout -s KILL 1200 python -m pytest tests/test_gpu_bigvgan.py tests/test_gpu_f5.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4 >> $OUT
cat $OUT

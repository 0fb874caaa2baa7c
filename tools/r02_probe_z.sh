#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
echo "== pytest ragged / pipeline" >> $OUT
timeout -s KILL 1200 python -m pytest tests/test_gpu_f5.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4 >> $OUT
echo "== bench config4" >> $OUT
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ['value','ms_per_step','dtype','clocks']}), json.dumps(d['e2e']['ms_per_step']), d['roofline']['frac'])
print({k:round(v) for k,v in d['profile_ms'].items() if v>25})
" >> $OUT 2>&1
cat $OUT

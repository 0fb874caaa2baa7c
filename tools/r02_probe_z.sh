#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
echo "== pytest -m gpu (all)" >> $OUT
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 >> $OUT
echo "== bench config4 --fp8 (level 2)" >> $OUT
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --fp8 2> gpurun_out/bench_z.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ['value','ms_per_step','dtype','clocks']}), json.dumps(d['e2e']['ms_per_step']), d['roofline']['frac'])
" >> $OUT 2>&1
cat $OUT

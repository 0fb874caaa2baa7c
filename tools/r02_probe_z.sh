#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
echo "== attn_time" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py attn_time 2>&1 | grep -v done >> $OUT
echo "== bvg_branches" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py bvg_branches 2>&1 | grep -v done >> $OUT
echo "== attention tests" >> $OUT
timeout -s KILL 300 python -m pytest tests/test_gpu_f5.py -m gpu -q -k "attention or chain" 2>&1 | tail -2 >> $OUT
cat $OUT

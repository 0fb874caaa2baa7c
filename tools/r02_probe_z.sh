#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
echo "== pytest f5 + fullsize" >> $OUT
timeout -s KILL 1200 python -m pytest tests/test_gpu_f5.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4 >> $OUT
echo "== f5_fp8 (timing part)" >> $OUT
timeout -s KILL 600 python tools/r02_probe.py f5_fp8 2>&1 | grep -E '"U"' >> $OUT
cat $OUT

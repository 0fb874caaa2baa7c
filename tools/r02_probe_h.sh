#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_l.log
: > $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_f5.py tests/test_gpu_fullsize.py -m gpu -x -q >> $OUT 2>&1
echo "rc=$?" >> $OUT
tail -30 $OUT

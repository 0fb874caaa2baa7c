#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-x}.log
: > $OUT
echo "== bvg_branches" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py bvg_branches >> $OUT 2>&1
echo "== pytest bigvgan / indextts / fullsize" >> $OUT
timeout -s KILL 900 python -m pytest tests/test_gpu_bigvgan.py tests/test_gpu_indextts.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8 >> $OUT
cat $OUT

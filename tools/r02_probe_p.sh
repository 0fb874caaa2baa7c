#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-p}.log
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 >> $OUT
echo "== ncu attention (new kernel)" >> $OUT
timeout -s KILL 600 ncu --set full --import-source on --clock-control none --kernel-name regex:attn_tc --launch-skip 30 --launch-count 1 -f \
  -o gpurun_out/ncu_attn_r02b python tools/f5_once.py 1 3 > gpurun_out/ncu_attn_r02b.log 2>&1
echo "rc=$?" >> $OUT
cat $OUT

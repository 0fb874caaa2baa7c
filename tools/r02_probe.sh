#!/bin/bash
# First GPU call of round 2: probe sections (one process each, short timeouts so that a protocol bug cannot hold the box),
# then the existing GPU test suite, then compute-sanitizer on the kernels VERDICT r01 names.
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-a}.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $OUT 2>&1
for sec in f5_small f5_full f5_time bigvgan; do
  echo "== section $sec" >> $OUT
  timeout -s KILL 300 python tools/r02_probe.py $sec >> $OUT 2>&1
  echo "rc=$?" >> $OUT
done
echo "== pytest -m gpu" >> $OUT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q >> $OUT 2>&1
echo "rc=$?" >> $OUT
tail -40 $OUT

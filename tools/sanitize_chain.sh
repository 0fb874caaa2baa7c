#!/bin/bash
# compute-sanitizer over the fused chain only (fp16 path + both e4m3 levels): tests/test_gpu_f5.py -k fused_chain_one_step
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitizer_${tool}_chain.log
  echo "== $tool chain: tests/test_gpu_f5.py -k fused_chain_one_step (fp16, e4m3 level 1 and 2)" > $log
  timeout -s KILL 600 env B200TTS_GRAPHS=0 $SAN --tool $tool --print-limit 20 --error-exitcode 9 \
      python -m pytest -m gpu -x -q tests/test_gpu_f5.py -k "fused_chain_one_step" >> $log 2>&1
  echo "rc=$?" >> $log
  tail -4 $log
done

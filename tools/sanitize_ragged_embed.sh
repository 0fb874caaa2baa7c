#!/bin/bash
# compute-sanitizer memcheck over the one-launch ragged embedding (cast_pad_rows_ragged_kernel + the 2*Ntot-row embedding GEMM):
# tests/test_gpu_f5.py -k ragged_embed, graphs off so every kernel is a plain launch
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
log=gpurun_out/sanitizer_memcheck_ragged_embed.log
echo "== memcheck: tests/test_gpu_f5.py -k ragged_embed (B200TTS_GRAPHS=0)" > $log
timeout -s KILL 170 env B200TTS_GRAPHS=0 $SAN --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest -m gpu -x -q tests/test_gpu_f5.py -k "ragged_embed" >> $log 2>&1
echo "rc=$?" >> $log
tail -6 $log

#!/bin/bash
# full ncu capture of the persistent decode kernel (one launch = up to 32 tokens) with source-level stall sampling
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gpt_decode_kernel' -s 1 -c 1 -f \
    -o gpurun_out/full_igpt_persist_$TAG python tools/prof_f5.py --what igpt --steps 40 > gpurun_out/full_igpt_persist_$TAG.log 2>&1
tail -3 gpurun_out/full_igpt_persist_$TAG.log
ls -la gpurun_out/*persist*

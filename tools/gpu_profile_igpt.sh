#!/bin/bash
# Run on the GPU box (gpurun): IndexTTS GPT decode -- bench lines, launch list, full capture of decode-step kernels.
#   tools/gpu_profile_igpt.sh TAG
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 300 python bench.py --workload indextts_gpt --steps 3 --warmup 3 2>gpurun_out/igpt_err_$TAG.log > gpurun_out/bench_igpt_bf16_$TAG.json
timeout 300 python bench.py --workload indextts_gpt --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/igpt_err_$TAG.log > gpurun_out/bench_igpt_f32_$TAG.json
# launch list: 2 sentences of 6 calls each (prefill + 5 decode steps; steps 3+ of the second sentence replay the graph)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_igpt_$TAG.csv \
    python tools/prof_f5.py --what igpt --steps 6 > gpurun_out/prof_igpt_$TAG.log 2>&1
# full capture: the gemv / attention / pick kernels of one decode step in the second sentence (skip the first sentence's launches)
timeout 400 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:'gemv_kernel|gpt_attn|gpt_pick' -s 400 -c 12 -f \
    -o gpurun_out/full_igpt_$TAG python tools/prof_f5.py --what igpt --steps 6 > gpurun_out/full_igpt_$TAG.log 2>&1
ls -la gpurun_out/ | tail -8
cut -c1-400 gpurun_out/bench_igpt_bf16_$TAG.json
cut -c1-300 gpurun_out/bench_igpt_f32_$TAG.json
tail -3 gpurun_out/igpt_err_$TAG.log

#!/bin/bash
# Run on the GPU box (gpurun): launch lists + full captures of the top kernels. Output under gpurun_out/.
#   tools/gpu_profile.sh TAG [lists] [full]
set -x
TAG=${1:-r01}
shift
WHAT="${*:-lists full}"
mkdir -p gpurun_out
if [[ "$WHAT" == *lists* ]]; then
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bigvgan_$TAG.csv \
    python tools/prof_f5.py --what bigvgan > gpurun_out/prof_bigvgan_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f5_$TAG.csv \
    python tools/prof_f5.py --what f5 --steps 2 > gpurun_out/prof_f5_$TAG.log 2>&1
fi
if [[ "$WHAT" == *full* ]]; then
# full captures (-s counts matching kernels only): the second call's first DiT layer (qkv GEMM, attention, out, ff1, ff2 ...),
# and BigVGAN stage-0 convs + AA activations of the second pass
ncu --set full --clock-control none --import-source on -k regex:'rowgemm_tc2|attn_tc' -s 230 -c 6 -f -o gpurun_out/full_f5_$TAG \
    python tools/prof_f5.py --what f5 --steps 2 > gpurun_out/full_f5_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rowgemm_tc2|aa_snake' -s 250 -c 6 -f -o gpurun_out/full_bigvgan_$TAG \
    python tools/prof_f5.py --what bigvgan > gpurun_out/full_bigvgan_$TAG.log 2>&1
fi
ls -la gpurun_out/ | tail -20

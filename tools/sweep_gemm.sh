#!/bin/bash
# Tile-shape sweep of rowgemm_tc on the hot-path shapes (GPU box). Output: gpurun_out/sweep_gemm_$1.log
TAG=${1:-x}
OUT=gpurun_out/sweep_gemm_$TAG.log
: > $OUT
echo "== auto" >> $OUT
python tools/bench_gemm.py >> $OUT 2>&1
for cfg in "256 128" "128 256" "256 256" "128 128" "256 192" "256 64" "128 64" "256 96"; do
  set -- $cfg
  echo "== BM=$1 BN=$2" >> $OUT
  B200TTS_BM=$1 B200TTS_BN=$2 python tools/bench_gemm.py dit vgan.s0 vgan.s1 vgan.s2 >> $OUT 2>&1
done

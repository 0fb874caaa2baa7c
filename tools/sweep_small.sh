#!/bin/bash
# small-C conv sweep (GPU box): A-ring depth and tile height. Output gpurun_out/sweep_small_$1.log
OUT=gpurun_out/sweep_small_${1:-x}.log
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python tools/bench_gemm.py vgan.s3 vgan.s4 vgan.s5 2>&1 | cut -c1-100 >> $OUT; }
run X=0
run B200TTS_NA=3
run B200TTS_NA=4
run B200TTS_BM=128 B200TTS_BN=96
run B200TTS_BM=128 B200TTS_BN=96 B200TTS_NA=4
run B200TTS_BM=128 B200TTS_BN=48 B200TTS_NA=4
run B200TTS_BM=128 B200TTS_BN=32 B200TTS_NA=4
run B200TTS_BM=256 B200TTS_BN=32 B200TTS_NA=4

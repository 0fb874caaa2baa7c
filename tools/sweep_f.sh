#!/bin/bash
OUT=gpurun_out/sweep_${1:-f}.log
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python tools/bench_gemm.py $FILTER 2>&1 | cut -c1-105 >> $OUT; }
FILTER=""
run X=0
FILTER="dit"
run B200TTS_KSPLIT=2
run B200TTS_KSPLIT=2 B200TTS_BM=128 B200TTS_BN=128
FILTER="vgan.s3 vgan.s4 vgan.s5"
run B200TTS_BM=512 B200TTS_BN=32
run B200TTS_BM=512 B200TTS_BN=48
run B200TTS_BM=128 B200TTS_BN=96

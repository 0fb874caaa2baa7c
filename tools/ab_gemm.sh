#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab_gemm.sh TAG
OUT=gpurun_out/ab_${1:-x}.log
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" python tools/bench_gemm.py $FILTER >> $OUT 2>&1; }
for lib in gpurun_ab/libb200tts_old.so text-to-speech-tts-onnx_b200/libb200tts.so; do
  FILTER="vgan.s0"; run B200TTS_LIB=$PWD/$lib B200TTS_BM=128 B200TTS_BN=256
  FILTER="vgan.s0"; run B200TTS_LIB=$PWD/$lib B200TTS_BM=256 B200TTS_BN=128
  FILTER="vgan.s1 vgan.s2"; run B200TTS_LIB=$PWD/$lib B200TTS_BM=128 B200TTS_BN=192
  FILTER="dit.out dit.ff2"; run B200TTS_LIB=$PWD/$lib B200TTS_BM=128 B200TTS_BN=128
  FILTER="dit.ff1 dit.qkv"; run B200TTS_LIB=$PWD/$lib B200TTS_BM=128 B200TTS_BN=256
done

#!/bin/bash
# A/B of gpurun_ab/lib_head.so (previous commit) against the working-tree build on ONE box: tools/ab_lib.sh TAG [bench_gemm filters...]
TAG=${1:-x}; shift
OUT=gpurun_out/ab_lib_$TAG.log
: > $OUT
for rep in 1 2; do
for lib in gpurun_ab/lib_head.so text-to-speech-tts-onnx_b200/libb200tts.so; do
  echo "== $lib (rep $rep)" >> $OUT
  B200TTS_LIB=$PWD/$lib python tools/bench_gemm.py "$@" 2>&1 | cut -c1-112 >> $OUT
  B200TTS_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bigvgan ms', round(d['ms_per_step'],3), {k.split('.')[-1]:round(v,2) for k,v in d['profile_ms'].items() if 'resconv' in k})" >> $OUT
done
done

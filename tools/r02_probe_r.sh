#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-r}.log
: > $OUT
for d in 15 0; do
  echo "== G=2 POLY=1 DBG=$d" >> $OUT
  B200TTS_ATTN_G=2 B200TTS_ATTN_POLY=1 B200TTS_ATTN_DBG=$d B200TTS_GRAPHS=0 B200TTS_ATTN_TRACE=gpurun_out/attn_trace.bin timeout -s KILL 100 python tools/attn_trace.py 2>&1 >> $OUT
done
cat $OUT

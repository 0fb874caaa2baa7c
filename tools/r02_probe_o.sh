#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-o}.log
: > $OUT
for p in 0; do
  echo "== determinism POLY=$p" >> $OUT
  B200TTS_ATTN_POLY=$p timeout -s KILL 300 python tools/attn_determinism.py >> $OUT 2>&1
done
cat $OUT

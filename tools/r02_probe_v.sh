#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-v}.log
: > $OUT
for u in 1 8; do
  echo "== chain trace U=$u" >> $OUT
  B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u$u.bin timeout -s KILL 300 python tools/chain_trace.py $u >> $OUT 2>&1
done
cat $OUT

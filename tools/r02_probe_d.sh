#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-e}.log
: > $OUT
for T in 8 4 2 1; do
  echo "== f5_small parity TEAM=$T" >> $OUT
  B200TTS_CHAIN_TEAM=$T timeout -s KILL 300 python tools/r02_probe.py f5_small 2>&1 | grep -E "chain\": 1|chain vs|one step|Error|error" >> $OUT
done
echo "== chain trace U=1" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u1.bin timeout -s KILL 200 python tools/chain_trace.py 1 >> $OUT 2>&1
echo "== chain trace U=8 (auto team)" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u8.bin timeout -s KILL 200 python tools/chain_trace.py 8 >> $OUT 2>&1
echo "== f5_time auto" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_time 2>&1 | grep -v profile_ms >> $OUT
for T in 2 4 8; do
  echo "== f5_time TEAM=$T" >> $OUT
  B200TTS_CHAIN_TEAM=$T timeout -s KILL 300 python tools/r02_probe.py f5_time 2>&1 | grep -E "chain\": 1" | grep -v profile_ms >> $OUT
done
echo "== f5_full parity (auto)" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_full >> $OUT 2>&1
cat $OUT

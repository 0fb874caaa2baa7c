#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-c}.log
: > $OUT
echo "== chain trace U=1" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u1.bin timeout -s KILL 200 python tools/chain_trace.py 1 >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== chain trace U=8" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u8.bin timeout -s KILL 200 python tools/chain_trace.py 8 >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== f5_time" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_time >> $OUT 2>&1
echo "== f5_small parity" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_small >> $OUT 2>&1
cat $OUT

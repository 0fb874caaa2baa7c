#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_b.log
: > $OUT
echo "== chain trace U=1" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u1.bin timeout -s KILL 200 python tools/chain_trace.py 1 >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== chain trace U=8" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u8.bin timeout -s KILL 200 python tools/chain_trace.py 8 >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== f5_time" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_time >> $OUT 2>&1
echo "== new tests" >> $OUT
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "fullsize or fp16 or chain or pipeline" >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== sanitizer" >> $OUT
SAN_TIMEOUT=400 TOOLS="memcheck racecheck" bash tools/sanitize.sh >> $OUT 2>&1
tail -100 $OUT

#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_k.log
: > $OUT
echo "== trace U=1" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u1.bin timeout -s KILL 200 python tools/chain_trace.py 1 2>&1 | grep -E "LN stat|slab published|epilogue tile done|MMA last" >> $OUT
echo "== trace U=8" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u8.bin timeout -s KILL 200 python tools/chain_trace.py 8 2>&1 | grep -E "LN stat|slab published|epilogue tile done|MMA last" >> $OUT
echo "== f5_time" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_time 2>&1 | grep '"chain": 1' | grep -v profile_ms >> $OUT
echo "== twins + small parity" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py twins f5_small 2>&1 | grep -E "twins|chain vs" >> $OUT
cat $OUT

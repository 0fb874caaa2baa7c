#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-f}.log
: > $OUT
echo "== chain trace U=1" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u1.bin timeout -s KILL 200 python tools/chain_trace.py 1 >> $OUT 2>&1
echo "== chain trace U=8 (auto team)" >> $OUT
B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace_u8.bin timeout -s KILL 200 python tools/chain_trace.py 8 >> $OUT 2>&1
echo "== f5_time auto" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_time 2>&1 | grep -v profile_ms >> $OUT
echo "== f5_small" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_small 2>&1 | grep -E "chain vs|one step|rror" >> $OUT
echo "== bench default" >> $OUT
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default_${1:-f}.json 2>> $OUT
echo "rc=$?" >> $OUT
echo "== pytest -m gpu" >> $OUT
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q >> $OUT 2>&1
echo "rc=$?" >> $OUT
cat $OUT | grep -vE "A producer reaches|kernel body|pdl_wait"

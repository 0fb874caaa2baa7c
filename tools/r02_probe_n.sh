#!/bin/bash
# Attention rewrite (P through tensor memory, shared row maximum, FMA-pipe exponentials): correctness for every POLY setting,
# time per F5 call, timeline.
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-n}.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $OUT 2>&1
for p in 0 1 2; do
  echo "== attention tests POLY=$p" >> $OUT
  B200TTS_ATTN_POLY=$p timeout -s KILL 300 python -m pytest tests/test_gpu_f5.py -m gpu -x -q -k attention >> $OUT 2>&1
  echo "rc=$?" >> $OUT
done
for p in 0 1 2; do
  echo "== attn_time POLY=$p" >> $OUT
  B200TTS_ATTN_POLY=$p timeout -s KILL 300 python tools/r02_probe.py attn_time >> $OUT 2>&1
  echo "rc=$?" >> $OUT
done
echo "== attention trace (POLY default)" >> $OUT
B200TTS_GRAPHS=0 B200TTS_ATTN_TRACE=gpurun_out/attn_trace.bin timeout -s KILL 200 python tools/attn_trace.py >> $OUT 2>&1
echo "== full-size parity" >> $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_f5.py -m gpu -x -q >> $OUT 2>&1
echo "rc=$?" >> $OUT
tail -60 $OUT

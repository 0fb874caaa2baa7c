#!/bin/bash
# Attention rewrite (P through tensor memory, shared row maximum, FMA-pipe exponentials, 2 or 4 softmax threads per row):
# correctness for every setting, time per F5 call, timeline.
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-n}.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $OUT 2>&1
for g in 2 4; do
for p in 0 1 2; do
  echo "== attention tests G=$g POLY=$p" >> $OUT
  B200TTS_ATTN_G=$g B200TTS_ATTN_POLY=$p timeout -s KILL 300 python -m pytest tests/test_gpu_f5.py -m gpu -x -q -k attention 2>&1 | tail -3 >> $OUT
  B200TTS_ATTN_G=$g B200TTS_ATTN_POLY=$p timeout -s KILL 300 python tools/attn_determinism.py 2>&1 | grep -v twins | head -2 >> $OUT
done
done
for g in 2 4; do
for p in 0 1 2; do
  echo "== attn_time G=$g POLY=$p" >> $OUT
  B200TTS_ATTN_G=$g B200TTS_ATTN_POLY=$p timeout -s KILL 300 python tools/r02_probe.py attn_time 2>&1 | grep -v done >> $OUT
done
done
echo "== attention trace (defaults)" >> $OUT
B200TTS_GRAPHS=0 B200TTS_ATTN_TRACE=gpurun_out/attn_trace.bin timeout -s KILL 200 python tools/attn_trace.py >> $OUT 2>&1
cat $OUT

#!/bin/bash
# LayerNorm folded into the chain's GEMMs: parity (full GPU suite), timing, twins
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-u}.log
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 >> $OUT
echo "== f5_full parity numbers" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py f5_full 2>&1 | cut -c1-300 >> $OUT
echo "== attn_time" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py attn_time 2>&1 >> $OUT
for t in 1 2 4 8; do
echo "== twins team $t" >> $OUT
B200TTS_CHAIN_TEAM=$t timeout -s KILL 300 python tools/r02_probe.py twins 2>&1 | cut -c1-300 >> $OUT
done
cat $OUT

#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_i.log
: > $OUT
# 2N mod 256 = offset of the second twin inside the 256-row blocks: 128 (other CTA of the pair), 32 (other TMEM quarter), 16, 2
for N in 1156 1153 1126; do TWIN_N=$N timeout -s KILL 200 python tools/r02_probe.py twins 2>&1 | grep '"chain": 1' >> $OUT; done
cat $OUT

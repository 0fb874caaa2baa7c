#!/bin/bash
# one `ncu --set full` capture of the fused DiT chain kernel (launch 40 of a config-3 utterance) with source correlation
mkdir -p gpurun_out
B200TTS_GRAPHS=0 timeout -s KILL 600 ncu --set full --import-source on --clock-control none --kernel-name regex:dit_chain \
  --launch-skip 40 --launch-count 1 -f -o gpurun_out/ncu_chain_${1:-a} python tools/chain_trace.py 1 > gpurun_out/ncu_chain_${1:-a}.log 2>&1
echo "rc=$?"
tail -5 gpurun_out/ncu_chain_${1:-a}.log
ls -la gpurun_out/*.ncu-rep

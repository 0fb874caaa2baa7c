#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe):
#  1. launch list (gpu__time_duration.sum per launch) of one config-3 utterance, fp16, fused chain: share of each kernel
#  2. `ncu --set full` of dit_chain_kernel, team of 8 (one utterance) and team of 1 (eight utterances), and of attn_tc_kernel
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f5_r02.csv \
  python tools/f5_once.py 1 > gpurun_out/launches_f5_r02.out 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/launches_f5_r02.out
timeout -s KILL 600 ncu --set full --import-source on --clock-control none --kernel-name regex:dit_chain --launch-skip 30 --launch-count 1 -f \
  -o gpurun_out/ncu_chain_team8_r02 python tools/f5_once.py 1 3 > gpurun_out/ncu_chain_team8_r02.log 2>&1
echo "chain team8 rc=$?"
timeout -s KILL 600 ncu --set full --import-source on --clock-control none --kernel-name regex:dit_chain --launch-skip 30 --launch-count 1 -f \
  -o gpurun_out/ncu_chain_team1_r02 python tools/f5_once.py 8 3 > gpurun_out/ncu_chain_team1_r02.log 2>&1
echo "chain team1 rc=$?"
timeout -s KILL 600 ncu --set full --import-source on --clock-control none --kernel-name regex:attn_tc --launch-skip 30 --launch-count 1 -f \
  -o gpurun_out/ncu_attn_r02 python tools/f5_once.py 1 3 > gpurun_out/ncu_attn_r02.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out/*r02*

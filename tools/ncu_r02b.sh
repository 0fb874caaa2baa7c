#!/bin/bash
# Round-2 final ncu evidence (B200_PROFILING.md recipe), after the LayerNorm fold / attention rewrite:
#  1. launch list (gpu__time_duration.sum per launch) of one config-3 utterance, fp16, fused chain: share of each kernel
#  2. `ncu --set full` of dit_chain_kernel inside the benchmark's own workload (configs[3], 64 ragged utterances, one launch),
#     team of 8 (one utterance), and of attn_tc_kernel (one utterance)
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f5_r02b.csv \
  python tools/f5_once.py 1 > gpurun_out/launches_f5_r02b.out 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/launches_f5_r02b.out
timeout -s KILL 900 ncu --set full --import-source on --clock-control none --kernel-name regex:dit_chain --launch-skip 40 --launch-count 1 -f \
  -o gpurun_out/ncu_chain_config4_r02b python bench.py --workload config4 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_chain_config4_r02b.log 2>&1
echo "chain config4 rc=$?"
timeout -s KILL 600 ncu --set full --import-source on --clock-control none --kernel-name regex:dit_chain --launch-skip 30 --launch-count 1 -f \
  -o gpurun_out/ncu_chain_team8_r02b python tools/f5_once.py 1 3 > gpurun_out/ncu_chain_team8_r02b.log 2>&1
echo "chain team8 rc=$?"
timeout -s KILL 600 ncu --set full --import-source on --clock-control none --kernel-name regex:attn_tc --launch-skip 30 --launch-count 1 -f \
  -o gpurun_out/ncu_attn_r02c python tools/f5_once.py 1 3 > gpurun_out/ncu_attn_r02c.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out/*r02b* gpurun_out/*r02c*

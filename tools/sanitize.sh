#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over the GPU tests that exercise the hand-synchronised kernels:
#   gpt_decode_kernel (barrier-free cross-CTA protocol), rowgemm_tc2sm_kernel (CTA pair, forced), attn_tc_kernel,
#   dit_chain_kernel (team hand-offs through global counters). Logs -> gpurun_out/sanitizer_<tool>_<subset>.log
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool subset env... -- pytest args
  local tool=$1 name=$2; shift 2
  local log=gpurun_out/sanitizer_${tool}_${name}.log
  echo "== $tool $name: $*" > $log
  timeout -s KILL ${SAN_TIMEOUT:-600} env B200TTS_GRAPHS=0 "${ENVV[@]}" $SAN --tool $tool --print-limit 20 --error-exitcode 9 \
      python -m pytest -m gpu -x -q "$@" >> $log 2>&1
  echo "rc=$?" >> $log
  tail -4 $log
}
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  ENVV=(B200TTS_2SM=1); run $tool conv2sm tests/test_gpu_bigvgan.py -k "conv1d_tcgen05_bf16 and (case0 or case5 or case9)"
  ENVV=(X=1); run $tool attention tests/test_gpu_f5.py -k "attention_tcgen05 and (130 or 257)"
  ENVV=(X=1); run $tool chain tests/test_gpu_f5.py -k "fused_chain_one_step"          # fp16 / bf16 operands and the e4m3 option
  ENVV=(X=1); run $tool gptdecode tests/test_gpu_indextts_gpt.py -k "bf16_generate_prefix or stop_token_in_the_middle"
done

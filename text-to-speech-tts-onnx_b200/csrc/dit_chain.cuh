// Fused row-block chain of one DiT block (dit_chain.cu):
//   x += gate_msa * (att @ Wout^T + b)  ->  LN-modulate  ->  GELU_tanh(. @ Wff1^T + b)  ->  x += gate_mlp * (. @ Wff2^T + b)
//   ->  LN-modulate (next block's attention modulation, or the final one)  ->  q|k|v of the NEXT block with RoPE / V^T epilogue
// as ONE persistent tcgen05 kernel. Reference semantics: F5_TTS/modeling_modified/F5/modules.py:599-613 (DiTBlock.forward),
// :301-305 (AdaLayerNorm), :329-340 (FeedForward), :459-466 (q/k/v + RoPE); dit.py:220 (norm_out).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "rowgemm_tc.cuh"

namespace b200tts {

struct DitChain {
  int R = 0;                       // rows = 2 * U * N (the CFG pair of every utterance in flight, stacked)
  int D = 0, FF = 0;               // model width (1024) and feed-forward width (2048)
  int f16 = 0;                     // 16-bit operand type: 0 = bf16, 1 = fp16
  int has_qkv = 1;                 // 0: last block -- stop after the final LN-modulate (n16b feeds proj_out)
  const __nv_bfloat16* att16 = nullptr;   // [R][D]  attention output (A operand of the out projection)
  float* x = nullptr;              // [R][D]  fp32 residual stream, updated in place
  __nv_bfloat16* n16 = nullptr;    // [R][D]  scratch: LN1-modulated rows (A operand of ff1)
  __nv_bfloat16* ff16 = nullptr;   // [R][FF] scratch: GELU(ff1)
  __nv_bfloat16* n16b = nullptr;   // [R][D]  LN2-modulated rows (A operand of the next q|k|v / of proj_out)
  const TcWeight *w_out = nullptr, *w_ff1 = nullptr, *w_ff2 = nullptr, *w_qkv = nullptr;
  const float *b_out = nullptr, *gate_msa = nullptr, *shift_mlp = nullptr, *scale_mlp = nullptr, *b_ff1 = nullptr, *b_ff2 = nullptr,
              *gate_mlp = nullptr, *shift_nxt = nullptr, *scale_nxt = nullptr, *b_qkv = nullptr;
  // q|k|v epilogue (rowgemm.cuh: rope_cs / vt_out)
  __nv_bfloat16* qk16 = nullptr;   // [R][2D] roped q | k
  const __half2* rope_cs = nullptr;
  int rope_rows = 1;
  const int2* rowinfo = nullptr;   // ragged batches: (sequence, position) per row (rowgemm.cuh)
  __nv_bfloat16* vt_out = nullptr;
  int vt_ld = 0, vt_heads = 0;
  // LayerNorm folded into the GEMMs: per-column vectors of THIS (block, Euler step), computed from the 16-bit weights
  //   u_ff1 = W_ff1 (1 + scale_mlp), v_ff1 = W_ff1 shift_mlp + b_ff1 [FF];  u_qkv / v_qkv likewise for the next block's q|k|v [3D]
  // and the per-row scale carried from LayerNorm to LayerNorm: rowscale[r] = 1 / std of row r at the LayerNorm in front of this
  // block's attention (written by ln_modulate for block 0 and by the previous chain launch afterwards; updated in place)
  const float *u_ff1 = nullptr, *v_ff1 = nullptr, *u_qkv = nullptr, *v_qkv = nullptr;
  float* rowscale = nullptr;
  // optional e4m3 mode of ff1 and q|k|v (tcgen05 kind::f8f6f4): weights [N][D] bytes with per-output-channel scales (u / v above
  // must then be computed from the DEQUANTISED weights); the A operands are emitted as e4m3 by the epilogues that produce x
  int fp8 = 0;                       // 1: ff1 + q|k|v; 2: ff2 as well (its A operand = the hidden activation written as 16 * GELU(.) in
                                     // e4m3 by ff1's epilogue; gate_mlp / b_ff2 must then carry the weight scale: gate * sw / 16, b * 16 / sw)
  const void *w8_ff1 = nullptr, *w8_qkv = nullptr, *w8_ff2 = nullptr;
  const float *sw_ff1 = nullptr, *sw_qkv = nullptr;
  // team synchronisation scratch (sizes below); `flags` must be zero when the kernel starts
  float* stats = nullptr;
  unsigned* flags = nullptr;
  unsigned long long* trace = nullptr;   // optional [CTAs][64] %globaltimer stamps of the hand-offs (debug: B200TTS_CHAIN_TRACE=<file>)
};

// The row-block schedule of one launch (dit_chain.cu, "two-phase"): phase 0 = `teams` teams of `team` CTA pairs walk row blocks
// [0, nrb0) in whole rounds; phase 1 = the `rem` remaining blocks, one team of `team1` pairs each. cost = modelled time in rounds at T = 1.
struct DitChainPlan { int team, teams, nrb0, rem, team1; double cost; };
DitChainPlan dit_chain_plan(int nrb, int pairs);
bool dit_chain_supported(int D, int FF, int H);      // shapes the kernel is specialised for + enough co-resident CTA pairs
size_t dit_chain_stats_floats(int R);
size_t dit_chain_flag_words(int R);
void dit_chain(const DitChain& c, cudaStream_t stream);

}  // namespace b200tts

// tcgen05 flash attention for the F5 DiT (attention_tc.cu).
#pragma once
#include "common.cuh"

namespace b200tts {

// S independent sequences of N tokens (F5: the CFG pair of each utterance of the batch, S = 2U)
// qk : bf16 [S][N][2*H*64] -- roped q in columns [0, H*64), roped k in [H*64, 2*H*64) (head h = 64-column group)
// vT : bf16 [S*H][64][ldv] -- V transposed per (sequence, head): vT[s*H + h][d][t]
// out: bf16 [S][N][H*64]   -- softmax(q k^T) v, heads concatenated (the layout the out-projection GEMM reads)
// f16 != 0: every 16-bit tensor (q, k, v, the probabilities, out) is IEEE fp16 instead of bf16
void attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vT, int ldv, __nv_bfloat16* out, int S, int N, int H, cudaStream_t stream, int f16 = 0);

}  // namespace b200tts

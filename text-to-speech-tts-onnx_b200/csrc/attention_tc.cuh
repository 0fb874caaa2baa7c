// tcgen05 flash attention for the F5 DiT (attention_tc.cu).
#pragma once
#include "common.cuh"

namespace b200tts {

// qk : bf16 [2][N][2*H*64] -- roped q in columns [0, H*64), roped k in [H*64, 2*H*64) (head h = 64-column group)
// vT : bf16 [2*H][64][ldv] -- V transposed per (batch, head): vT[b*H + h][d][t]
// out: bf16 [2][N][H*64]   -- softmax(q k^T) v, heads concatenated (the layout the out-projection GEMM reads)
void attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vT, int ldv, __nv_bfloat16* out, int N, int H, cudaStream_t stream);

}  // namespace b200tts

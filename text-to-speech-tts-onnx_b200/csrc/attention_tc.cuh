// tcgen05 flash attention for the F5 DiT (attention_tc.cu).
#pragma once
#include "common.cuh"

namespace b200tts {

// S independent sequences of N tokens (F5: the CFG pair of each utterance of the batch, S = 2U)
// qk : bf16 [S][N][2*H*64] -- roped q in columns [0, H*64), roped k in [H*64, 2*H*64) (head h = 64-column group)
// vT : bf16 [S*H][64][ldv] -- V transposed per (sequence, head): vT[s*H + h][d][t]
// out: bf16 [S][N][H*64]   -- softmax(q k^T) v, heads concatenated (the layout the out-projection GEMM reads)
// f16 != 0: every 16-bit tensor (q, k, v, the probabilities, out) is IEEE fp16 instead of bf16
// Ragged batch (d_seq_off / d_seq_len non-null, device int[S]): the sequences are concatenated, qk / out are [total_rows][..]
// with sequence s at rows [seq_off[s], seq_off[s] + seq_len[s]); N = the longest sequence; V^T keeps ldv columns per (sequence,
// head) and must hold finite values beyond a sequence's length (they meet probabilities of exactly zero).
void attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vT, int ldv, __nv_bfloat16* out, int S, int N, int H, cudaStream_t stream, int f16 = 0,
                  const int* d_seq_off = nullptr, const int* d_seq_len = nullptr, long total_rows = 0);

}  // namespace b200tts

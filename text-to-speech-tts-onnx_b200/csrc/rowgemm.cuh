// "Shifted-row GEMM": the one dense contraction of the whole hot path, on channels-last tensors.
//
//   out[b, t, g*N + n] = epi( sum_{j<taps} sum_{c<Cin} X[b, t + (j-center)*dil, g*Cin + c] * W[g][j][c][n] )
//
// with X rows outside [0, Lin) reading as zero. It covers
//   * Linear layers                     taps=1                      (DiT, Vocos pointwise, mel fbank)
//   * dilated "same" Conv1d             taps=k, center=(k-1)/2      (BigVGAN resblocks, conv_pre, Vocos embed)
//   * grouped Conv1d                    groups=16                   (DiT conv position embedding)
//   * ConvTranspose1d stride u, k=2u    taps=2, center=1, N=u*Cout, flat output shift -(u/2)*Cout
//                                                                   (BigVGAN upsamplers, see bigvgan.cu)
//   * STFT / ISTFT bases                taps=1 with ldx = hop (overlapping rows)
// Two implementations share this parameter block: rowgemm_f32 (SIMT fp32, the parity engine) and
// rowgemm_tc (tcgen05 + TMA, bf16 operands / fp32 accumulate, the fast engine).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace b200tts {

struct RowGemm {
  // A operand (activations), fp32 for rowgemm_f32, bf16 for rowgemm_tc
  const void* x = nullptr;
  long x_bstride = 0;     // elements between batches
  int ldx = 0;            // elements between rows
  int Lin = 0;            // valid input rows per batch: [0, Lin)
  // B operand (weights): f32 path W[g][j][c][n] (n contiguous, row stride ldw);
  //                      tc  path W[g][j][n][c] bf16 (c contiguous), described by a tensor map
  const void* w = nullptr;
  int ldw = 0;
  int Cin = 0, N = 0, taps = 1, dil = 1, center = 0, groups = 1;
  int M = 0;              // output rows per batch
  int B = 1;
  // output: flat index = b*o_bstride + t*ldo + g*N + n + o_shift, written iff 0 <= (t*ldo + g*N + n + o_shift) < o_limit
  void* out = nullptr;
  long o_bstride = 0;
  int ldo = 0;
  long o_shift = 0;
  long o_limit = 0;       // 0 -> M*ldo
  int out_bf16 = 0;       // type code of `out`: 0 = fp32, 1 = bf16, 2 = fp16
  int f16 = 0;            // tensor-core path: operand (A, W, out2, vt_out) 16-bit type: 0 = bf16, 1 = fp16
  __nv_bfloat16* out2 = nullptr;   // tensor-core path only: optional second, bf16, copy of the output (same indexing)
  // epilogue: v = acc + bias[n]; v = act(v); v *= gate[n]; v += res[idx]; if (accumulate) v += out[idx]; v *= scale
  const float* bias = nullptr;    // [groups*N]
  const float* gate = nullptr;    // [groups*N]
  const float* res = nullptr;     // indexed like out (fp32)
  int accumulate = 0;
  float scale = 1.0f;
  int act = ACT_NONE;
  // tensor-core path only: fused q/k/v epilogue of the DiT attention (F5 modules.py:459-466).
  //   columns [0, rope_cols): interleaved-pair RoPE with the packed table rope_cs[rope_rows][64] of (cos, sin) fp16 pairs
  //                           (exact: the reference rounds its tables through fp16, Export_F5.py:111-112) at t = row % rope_rows
  //   columns >= vt_col0    : written transposed, vt_out[((row / rope_rows) * heads + h) * 64 + d][t] (row stride vt_ld)
  const __half2* rope_cs = nullptr;
  int rope_cols = 0, rope_rows = 1;
  const int2* rowinfo = nullptr;   // ragged batches: (sequence, position) per output row instead of row / rope_rows, row % rope_rows
  __nv_bfloat16* vt_out = nullptr;
  int vt_col0 = 0, vt_ld = 0, vt_heads = 0;
};

// SIMT fp32 implementation (rowgemm_f32.cu). Requires Cin % 4 == 0, N % 4 == 0, ldx/ldw/ldo % 4 == 0.
void rowgemm_f32(const RowGemm& p, cudaStream_t stream);

}  // namespace b200tts

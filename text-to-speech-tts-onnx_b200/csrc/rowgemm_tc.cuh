// tcgen05 + TMA implementation of the shifted-row GEMM (rowgemm.cuh): bf16 operands, fp32 accumulation
// in TMEM. See rowgemm_tc.cu for the kernel.
#pragma once
#include <cuda.h>

#include "engine.cuh"
#include "rowgemm.cuh"

namespace b200tts {

// Weights for the tensor-core path: bf16 W[g*taps + j][n][c] (c contiguous, row stride ldc). The TMA map over it is
// encoded per launch, because its box height is the N tile the launch picks for the problem's M (rowgemm_tc.cu: pick_tile).
struct TcWeight {
  DevBuf<__nv_bfloat16> w;
  int Cin = 0, ldc = 0, N = 0, taps = 0, groups = 1;
  int f16 = 0;            // 16-bit type of the stored weights: 0 = bf16, 1 = fp16
  bool ready = false;
};

// Build from an fp32 device tensor already laid out as [groups*taps][N][Cin] (c contiguous).
void tc_weight_from_f32(TcWeight& tw, const float* w_gjnc, int groups, int taps, int N, int Cin, cudaStream_t s, int f16 = 0);

// A operand: bf16, rows of p.ldx elements (ldx % 8 == 0). Output fp32 or bf16 (p.out_bf16).
void rowgemm_tc(const RowGemm& p, const TcWeight& w, cudaStream_t stream);

// bf16 3-D TMA map: dims {d0 (contiguous), d1, d2}, element strides {ld1, ld2}, box {64, box1, 1}, SWIZZLE_128B, zero OOB fill
void tc_encode_map(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2,
                   uint32_t box1);

// f16 != 0: the 16-bit side is IEEE fp16 instead of bf16 (same storage type)
// 2-D TMA map for epilogue tiles: dims {d0 (contiguous), d1} of `elem_bytes`-wide elements (2 or 4), row stride ld1 elements,
// box {box0, box1} with box0 * elem_bytes = 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B); rows / columns beyond the dims are
// zero-filled on load and clipped on store
// 8-bit (e4m3) operand map: dims {d0 bytes, d1, d2}, byte strides {ld1, ld2}, box {128, box1, 1}, SWIZZLE_128B, zero OOB fill
void tc_encode_map_u8(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2, uint32_t box1);
void tc_encode_map2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t ld1, uint32_t box0,
                     uint32_t box1);

void cast_f32_to_bf16(const float* in, __nv_bfloat16* out, long n, cudaStream_t s, int f16 = 0);
void cast_bf16_to_f32(const __nv_bfloat16* in, float* out, long n, cudaStream_t s, int f16 = 0);
// (rows, C) fp32 -> (rows, ldo) bf16 with zero-filled padding columns
void cast_pad_f32_to_bf16(const float* in, __nv_bfloat16* out, long rows, int C, int ldo, cudaStream_t s, int f16 = 0);
// out[r][:] = in[token of DiT row r][:] (zero-padded to ldo) for a ragged batch: rowinfo[r] = (sequence, position), seq_off[sequence] =
// first DiT row of the sequence; the two CFG sequences 2u, 2u+1 of utterance u read the same token rows seq_off[2u] / 2 + position
void cast_pad_rows_ragged(const float* in, __nv_bfloat16* out, const int2* rowinfo, const int* seq_off, long rows, int C, int ldo,
                          cudaStream_t s, int f16 = 0);

}  // namespace b200tts

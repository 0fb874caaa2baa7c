// Small layout helpers (reference (B, C, L) <-> engine channels-last (B, L, C)).
#pragma once
#include "common.cuh"

namespace b200tts {

// in (B, R, C) -> out (B, C, R), fp32
void batched_transpose(const float* in, float* out, int B, int R, int C, cudaStream_t s);

// Conv1d weight (Cout, Cin/groups, k) fp32 -> [g][j][c][n] (n_major = 1: the SIMT rowgemm layout, n contiguous)
//                                          or [g][j][n][c] (n_major = 0: the tensor-core layout, c contiguous)
void conv_weight_permute(const float* W, float* out, int Cout, int cin_g, int k, int groups, int n_major, cudaStream_t s);
// in [R][C] -> out [C][ldo] (ldo >= R, padding columns zero)
void transpose_pad(const float* in, float* out, int R, int C, int ldo, cudaStream_t s);

}  // namespace b200tts

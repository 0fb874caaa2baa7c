// Small layout helpers (reference (B, C, L) <-> engine channels-last (B, L, C)).
#pragma once
#include "common.cuh"

namespace b200tts {

// in (B, R, C) -> out (B, C, R), fp32
void batched_transpose(const float* in, float* out, int B, int R, int C, cudaStream_t s);

}  // namespace b200tts

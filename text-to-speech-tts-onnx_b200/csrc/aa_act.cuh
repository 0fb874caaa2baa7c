// Fused anti-aliased SnakeBeta activation (aa_act.cu).
#pragma once
#include "common.cuh"

namespace b200tts {

// Upload the 12 Kaiser-sinc taps (computed on the host exactly as filter.py:30-62 does).
void aa_set_filter(const float* taps12_host);

// x (B, L, C) -> y (B, L, C)  [post: y (B, L+30, C)], channels-last. alpha = exp(alpha_log),
// inv_beta = 1/(exp(beta_log)+1e-9), both [C]. precise: sinf (fp32 parity mode) vs __sinf.
void aa_snake(const void* x, int in_bf16, void* y, int out_bf16, const float* alpha, const float* inv_beta,
              int B, int C, int L, bool precise, bool post, cudaStream_t stream);

}  // namespace b200tts

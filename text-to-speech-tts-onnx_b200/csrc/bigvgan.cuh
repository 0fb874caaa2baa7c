// BigVGAN-v2 generator (mel -> int16 PCM) on channels-last tensors. See bigvgan.cu.
#pragma once
#include <string>

#include "engine.cuh"

namespace b200tts {

struct BigVGANModel;

// Build device-side layouts from engine.weights["<prefix>*"] (reference state_dict names). prefix "bigvgan." = the
// mel -> PCM session (BigVGAN/Export_BigVGAN.py); "ivgan." = the vocoder half of IndexTTS_F (IndexTTS/Export_IndexTTS.py:
// 292-314, IndexTTS/modeling_modified/models.py:130-250): same generator with a LayerNorm'd GPT latent as input, a
// conditioning vector added after conv_pre and after every upsampler, kernel = stride upsamplers and a conv_post bias.
BigVGANModel* bigvgan_build(Engine& e, const std::string& prefix);

// Mel vocoder (d_conds == nullptr): d_in = (B, n_mels, T) fp32 device (reference layout, Export_BigVGAN.py:65-70).
// IndexTTS (d_conds = nstages + 1 device vectors: cond_0..cond_{n-1} of C_i floats, then the cond_layer vector of C0):
//   d_in = latent rows (T, gpt_dim) fp32 device = hidden[:-2], B = 1.
// d_pcm: (B, hop*T+30) int16 device; d_wave (optional): same shape fp32, the pre-cast value tanh(.)*32767 clamped.
// cl_bstride > 0 (mel vocoder only): d_in is ALREADY channels-last, (B, T, n_mels) rows with cl_bstride floats between batches --
// how the F5 DiT leaves its mel (f5.cu: noise [U][N][n_mels], generated frames = rows ref_len..N-1), so the F5 -> BigVGAN
// pipeline hands the frames over without the (B, n_mels, T) round trip.
void bigvgan_forward(Engine& e, BigVGANModel& m, const float* d_in, int B, int T, int precision, int16_t* d_pcm, float* d_wave,
                     const float* const* d_conds = nullptr, long cl_bstride = 0);

void bigvgan_free(BigVGANModel* m);
int bigvgan_num_mels(const BigVGANModel& m);        // conv_pre input width (n_mels, or gpt_dim for IndexTTS)
int bigvgan_num_stages(const BigVGANModel& m);
int bigvgan_stage_channels(const BigVGANModel& m, int i);   // i = -1: conv_pre output channels
long bigvgan_out_samples(const BigVGANModel& m, int T);   // hop*T + 30

// 16-bit weight layouts for the tensor-core path (idempotent per operand type: f16 = 0 bf16, 1 fp16).
void bigvgan_tc_prepare(Engine& e, BigVGANModel& m, int f16 = 0);

}  // namespace b200tts

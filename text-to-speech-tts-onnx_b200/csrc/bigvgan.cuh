// BigVGAN-v2 generator (mel -> int16 PCM) on channels-last tensors. See bigvgan.cu.
#pragma once
#include "engine.cuh"

namespace b200tts {

struct BigVGANModel;

// Build device-side layouts from engine.weights["bigvgan.*"] (reference state_dict names).
BigVGANModel* bigvgan_build(Engine& e);

// d_mel: (B, n_mels, T) fp32 device (reference layout, BigVGAN/Export_BigVGAN.py:65-70);
// d_pcm: (B, 256*T+30) int16 device; d_wave (optional): same shape fp32, the pre-cast value
// tanh(.)*32767 clamped (for tolerance analysis in tests).
void bigvgan_forward(Engine& e, const float* d_mel, int B, int T, int precision, int16_t* d_pcm, float* d_wave);

void bigvgan_free(BigVGANModel* m);
int bigvgan_num_mels(const Engine& e);
long bigvgan_out_samples(const Engine& e, int T);   // hop*T + 30

// bf16 weight layouts + TMA maps for the tensor-core path (idempotent).
void bigvgan_tc_prepare(Engine& e);

}  // namespace b200tts

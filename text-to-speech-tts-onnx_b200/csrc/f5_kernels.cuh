// Element-wise / reduction kernels of the F5-TTS graphs (everything that is not a GEMM or attention).
// All tensors are row-major (rows = time, columns = channels), fp32 unless noted.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace b200tts {

// out = LayerNorm(x; no affine, eps 1e-6) * (1 + scale) + shift   (F5 modules.py:301-305,321-325,609)
// x [R][D] fp32; scale, shift [D]; out fp32 (out_bf16 = 0), bf16 (1) or fp16 (2) with row stride D.
void ln_modulate(const float* x, const float* scale, const float* shift, void* out, int out_bf16, int R, int D, cudaStream_t s,
                 float* rstd_out = nullptr /* optional [R]: 1 / std of every row */);
// u[t][j] = sum_k W16[j][k] (1 + scale[t][k]), v[t][j] = sum_k W16[j][k] shift[t][k] + bias[j], t < nfe (rows of a modulation table
// with row stride mod_ld); W16 = the [N][ldc] 16-bit tensor-core weight (K <= 1024 columns used)
// f16 = 2: W16 is an e4m3 byte tensor with per-row scales `wscale` (quantize_rows_e4m3)
void fold_vectors(const void* w16, int ldc, int K, int f16, const float* scale, const float* shift, int mod_ld, const float* bias, float* u,
                  float* v, int N, int nfe, cudaStream_t s, const float* wscale = nullptr);
// per-output-channel e4m3 quantisation of an fp32 weight [N][K]: q [N][ldq] bytes, scale [N] (W ~ q * scale)
void quantize_rows_e4m3(const float* w, int N, int K, void* q, int ldq, float* scale, cudaStream_t s);
// gate8[t][j] = gate[t][j] * sw[j] / hgain (t < nfe, gate rows mod_ld apart), bias8[j] = bias[j] * hgain / sw[j]
void fold_gate_bias(const float* gate, int mod_ld, const float* bias, const float* sw, float hgain, float* gate8, float* bias8, int N, int nfe,
                    cudaStream_t s);
// nn.LayerNorm(D, eps) with affine (text ConvNeXtV2 block, modules.py:248)
void layernorm_affine(const float* x, const float* w, const float* b, float* out, int R, int D, float eps, cudaStream_t s);
void layernorm_affine_bf16(const float* x, const float* w, const float* b, __nv_bfloat16* out, int R, int D, float eps, cudaStream_t s);
// Vocos "LayerNorm": w * x / ||x||_2 + b per row, w already x sqrt(C) (vocos/models.py:80,83; modules.py:46)
void l2_norm_affine(const float* x, const float* w, const float* b, float* out, int R, int C, cudaStream_t s);
// depthwise Conv1d k=7 pad 3 on (B, L, C): w [7][C], bias [C]
void dwconv7(const float* x, const float* w, const float* bias, float* out, int B, int L, int C, cudaStream_t s);
// GRN (modules.py:217-226) on one sequence x [R][C] in place: gamma*(x*Nx)+beta+x with Gx = ||x||_2 over ROWS
void grn_inplace(float* x, const float* gamma, const float* beta, float* scratch /* >= C + 1 */, int R, int C, cudaStream_t s);
// text embedding front (dit.py:49-63): ids [N] (0 = filler) -> out [N][D] = (table[use_ids ? ids : 0] + pos[n]) masked by ids == 0
void text_embed_gather(const int* ids, const float* table, const float* pos, float* out, int N, int D, int use_ids, cudaStream_t s);
void mask_rows(float* x, const int* ids, int N, int D, cudaStream_t s);   // rows with ids == 0 -> 0
// ids_out[n] = n < n_text ? text_ids[n] + 1 : 0   (Export_F5.py:136)
void pad_text_ids(const int* text_ids, int n_text, int* ids_out, int N, cudaStream_t s);
// audio int16 [L] -> float / 32768 with reflect padding of `pad` on both sides: out [L + 2*pad] (STFT_Process.py:144-147)
void audio_reflect_pad(const int16_t* audio, float* out, long L, int pad, cudaStream_t s);
// spec [F][ld] (real at cols [0,bins), imag at [bins, 2*bins)) -> mag [F][ldm] = sqrt(re^2+im^2), padding cols zero
void stft_magnitude(const float* spec, int ld, float* mag, int ldm, int F, int bins, cudaStream_t s);
// dst[n][col0 + c] = n < F ? log(max(mel[n][c], 1e-5)) : 0 ; rows n < N, c < C  (Export_F5.py:125-130)
void logmel_into(const float* mel, int F, float* dst, int ld_dst, int col0, int N, int C, cudaStream_t s);
// dst[n][col0 + c] = src[n][c] (copy a column block), or zeros when src == nullptr
void copy_cols(const float* src, int ld_src, float* dst, int ld_dst, int col0, int N, int C, cudaStream_t s);
// Euler + CFG (Export_F5.py:179-181): noise += (p0 + (p0 - p1) * cfg) * dt ; pred [U][2][n], noise [U][n]
void euler_cfg_update(float* noise, const float* pred, long n, int U, float cfg, float dt, cudaStream_t s);
// head [G][ld] (log-mag cols [0,bins), phase cols [bins, 2 bins)) -> out [G][ld] = [min(exp(m),100)*cos p | ..*sin p | 0]
void istft_input(const float* head, float* out, int G, int bins, int ld, cudaStream_t s);
// frames [G][nfft] -> pcm [hop*(G-1)]: overlap-add, crop nfft/2 both sides, * window_sum_inv, clamp +-1, *32767, truncate
// (STFT_Process.py:160-166, Export_F5.py:201-203); wave (optional) gets the pre-cast float
void istft_overlap_add(const float* frames, const float* window_sum_inv, int G, int nfft, int hop, int16_t* pcm, float* wave, cudaStream_t s);
// fp32 attention helpers (parity engine)
// qkv [R][3*D] fp32 (R = 2N) -> q (roped, written back in place into cols [0,D)), kT [2][H][hd][ldk], v [2][H][ldk][hd]
void rope_split_f32(float* qkv, const float* cos, const float* sin, float* kT, float* v, int N, int H, int hd, int ldk, cudaStream_t s);
// softmax over the first n of ld columns of every row (rest set to 0), in place
void softmax_rows(float* x, long rows, int n, int ld, cudaStream_t s);
// out[i] = (cos[i], sin[i]) as an fp16 pair (the RoPE tables are fp16-rounded by the export, Export_F5.py:111-112: exact)
void rope_pack_half(const float* cos, const float* sin, __half2* out, long n, cudaStream_t s);
// y = silu(x)
void silu(const float* x, float* y, long n, cudaStream_t s);

}  // namespace b200tts

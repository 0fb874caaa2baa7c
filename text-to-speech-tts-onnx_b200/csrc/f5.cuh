// F5-TTS graphs on the GPU: F5_Preprocess (STFT -> log-mel, text embedding), F5_Transformer (DiT step on the CFG
// pair + Euler update), F5_Decode (Vocos + ISTFT -> int16). See f5.cu.
#pragma once
#include "engine.cuh"

namespace b200tts {

struct F5Model;

F5Model* f5_build(Engine& e);
void f5_free(F5Model* m);

// Per-utterance device state (owned by the model; one utterance in flight per engine).
// Graph A (Export_F5.py:117-141): audio int16 [L] (device), text_ids int32 [n_text] (device), N = max_duration.
// Fills the state's cond / cond_drop [N][612], ref_signal_len, rope rows. `noise` is NOT drawn here: the caller
// supplies it through f5_set_noise (the reference draws it with ORT's RandomNormalLike, irreproducible elsewhere).
// u / U: slot of this utterance in a batch of U utterances that share N (length-bucketed batching; default: one utterance).
// fast: the text-embedding GEMMs run on tensor cores (bf16 operands); the per-graph session entry points keep fp32.
void f5_preprocess(Engine& e, const int16_t* d_audio, long L, const int* d_text_ids, int n_text, int N, int u = 0, int U = 1,
                   int fast = 0 /* engine.cuh Precision: 0 fp32 SIMT front end, 1 / 2 bf16 / fp16 tensor-core GEMMs */);
int f5_ref_len(const Engine& e);
int f5_seq_len(const Engine& e);
int f5_cond_dim(const Engine& e);   // n_mels + text_dim (612)
int f5_n_mels(const Engine& e);
int f5_nfe(const Engine& e);
// device pointers of the current utterance's tensors (fp32): cond / cond_drop [N][612], noise [N][100]
float* f5_cond(Engine& e, int u = 0);
float* f5_cond_drop(Engine& e, int u = 0);
float* f5_noise(Engine& e, int u = 0);
// Set up state for externally supplied graph-B inputs (the per-step session path): allocates for N rows.
void f5_begin(Engine& e, int N, int U = 1);
// Ragged batch: utterance u has Ns[u] frames. Lays out the concatenated tensors and uploads the per-sequence /
// per-row tables (synchronises: call it outside a stream capture); f5_preprocess(.., Ns[u], u, U) then fills utterance u.
void f5_begin_ragged(Engine& e, int U, const int* Ns);
void f5_restore_ragged(Engine& e, int U, const int* Ns, const int* Fs);
long f5_tok_offset(const Engine& e, int u);     // rows of noise / cond before utterance u
// rope rows [N][64] fp32 device (cos, sin) -- defaults to the model's fp16-rounded tables, may be overridden
void f5_set_rope(Engine& e, const float* d_cos, const float* d_sin);
// same, but returns the device buffers ([N][64] each) for the caller to fill (then selects them)
void f5_rope_buffers(Engine& e, float** d_cos, float** d_sin);
// host-side shape state of the current utterance (what f5_preprocess sets), for graph replays that skip the enqueue code
void f5_restore_shape(Engine& e, int N, int ref_len, int U = 1);
// must be called after cond / cond_drop changed and before f5_steps: precomputes the step-invariant half of the
// input embedding (W_c . cond + b for both CFG rows)
void f5_prepare_cond(Engine& e);
// Graph B (Export_F5.py:177-182) `count` times starting at time_step `first`; noise updated in place.
void f5_steps(Engine& e, int first, int count, int precision);
// Graph C (Export_F5.py:197-203): decode noise[ref_len:] -> pcm int16 [256 * (N - ref_len - 1)] (device);
// d_mel (N x 100 fp32 device) defaults to the state's noise when null.
// fast: Vocos / ISTFT GEMMs on tensor cores (bf16 operands, fp32 accumulate and residual stream).
long f5_decode(Engine& e, const float* d_mel, int N, int ref_len, int16_t* d_pcm, float* d_wave, int fast = 0);

}  // namespace b200tts

/* libb200tts -- C ABI of the B200-native engine for the F5-TTS / BigVGAN hot path of
 * DakeQQ/Text-to-Speech-TTS-ONNX.
 *
 * The reference has no FFI: its boundary is onnxruntime.InferenceSession(path).run(names, feed) called from
 * Python (F5_TTS/F5-TTS-ONNX-Inference.py:173,193,214,247,292,306; BigVGAN/Export_BigVGAN.py:153,170).
 * Each entry point below replaces one such session (or the loading of its initializers) and is what the
 * Python shim in text-to-speech-tts-onnx_b200/session.py binds with ctypes.
 * Plain C types only; every function returns 0 on success and a non-zero code on failure, with the message
 * available from b200tts_last_error() (thread-local). There is no CPU fallback: without a CUDA device
 * b200tts_create fails.
 *
 * Tensors cross the boundary in the reference's own layouts and dtypes (Export_F5.py:294-306,354-365,409-414;
 * Export_BigVGAN.py:65-70). "host" pointers are ordinary (ideally pinned) host memory; "_device" variants take
 * device pointers on the engine's device and enqueue on the engine's stream without synchronising.
 */
#ifndef B200TTS_H
#define B200TTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200tts_engine b200tts_engine;

/* arithmetic of the dense contractions: fp32 on CUDA cores (the parity engine), or bf16 / IEEE fp16 operands on the tcgen05
 * tensor cores with fp32 accumulation. BASELINE.json's configurations name fp16 (the reference's own recipe:
 * F5_TTS/Export_F5.py:139-141,321-333); the IndexTTS GPT-2 entry points take F32 or BF16 only. */
enum { B200TTS_F32 = 0, B200TTS_BF16 = 1, B200TTS_F16 = 2 };

/* ---- lifetime / errors ------------------------------------------------------------------------------ */
/* Replaces InferenceSession construction (F5-TTS-ONNX-Inference.py:173-214): one engine per GPU. */
int b200tts_create(int device, b200tts_engine** out);
void b200tts_destroy(b200tts_engine* e);
const char* b200tts_last_error(void);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the engine's own stream. */
int b200tts_set_stream(b200tts_engine* e, void* cuda_stream);
int b200tts_synchronize(b200tts_engine* e);
/* Engine switches (the role of the reference's provider / session options, F5-TTS-ONNX-Inference.py:41-85,152-169):
 *   "dit_chain"   1 (default): every F5 DiT block runs as attention + ONE fused row-block kernel; 0: seven launches per block
 *   "cuda_graphs" 1 (default): repeated calls of one shape replay a captured CUDA graph; 0: enqueue kernel by kernel
 *   "bigvgan_branches" 1 (default): the three resblocks of a BigVGAN stage run as concurrent branches; 0: one after the other
 *   "ragged_embed" 1 (default): the input embedding of a ragged batch is one launch over all DiT rows; 0: two launches per utterance
 *   "dit_fp8"     0 (default); 1: ff1 and q|k|v of the fused chain take e4m3 operands (tcgen05 kind::f8f6f4); 2: ff2 as well.
 *                 Lower-fidelity modes (PCM SNR ~32 / ~30.5 dB against the fp32 reference instead of ~62 dB for fp16), 9 / 15 %
 *                 faster on eight utterances
 * Results do not depend on cuda_graphs or bigvgan_branches (bit-identical) nor on ragged_embed (same dot products per row); dit_chain changes how the LayerNorm is applied
 * (folded into the GEMM epilogues, same tolerance class). */
int b200tts_set_option(b200tts_engine* e, const char* name, int value);
/* Number of kernels this library has launched so far in this process (bench.py: gpu_launches). */
unsigned long long b200tts_launch_count(void);
/* Host arithmetic only (no GPU needed; unit-tested on CPU): the row-block schedule the fused DiT chain would use for `row_blocks`
 * blocks of 256 rows on `resident_pairs` CTA pairs -> plan5 = {team, teams, blocks in phase 0, remaining blocks, team of phase 1}. */
int b200tts_debug_chain_plan(int row_blocks, int resident_pairs, int* plan5);

/* ---- weights (replace the .onnx initializers written by Export_F5.py / Export_BigVGAN.py) ------------ */
/* name = "<model>.<reference state_dict name>", model in {bigvgan, dit, vocos}; fp32, row-major, rank <= 4.
 * The export-time transforms (Export_F5.py:321-333,389-402; Export_BigVGAN.py:53-57) are applied by the caller
 * (weights.py) exactly where the reference applies them. */
int b200tts_load_tensor(b200tts_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim);
int b200tts_load_tensor_device(b200tts_engine* e, const char* name, const float* dev_data, const int64_t* shape, int ndim);
/* Build the device layouts of one model after its tensors are loaded. */
int b200tts_bigvgan_build(b200tts_engine* e);

/* ---- BigVGAN session (BigVGAN/Export_BigVGAN.py:44-49,65-70,170) ----------------------------------------
 * mel_features (B, n_mels, T) fp32 -> generated_wav (B, 1, 256*T+30) int16. wave_out (optional, may be NULL)
 * receives the pre-cast float value tanh(.)*32767 clamped to [-32768, 32767]. */
int b200tts_bigvgan_run(b200tts_engine* e, const float* mel_host, int B, int T, int precision,
                        int16_t* pcm_host, float* wave_host);
int b200tts_bigvgan_run_device(b200tts_engine* e, const float* mel_dev, int B, int T, int precision,
                               int16_t* pcm_dev, float* wave_dev);

/* ---- IndexTTS_F session: the vocoder half of IndexTTS (IndexTTS/Export_IndexTTS.py:292-314, run at
 * IndexTTS/Inference_IndexTTS_ONNX.py:787) -- generality check of the BigVGAN kernels on a second model (config 5).
 * Tensors "ivgan.<IndexTTS bigvgan state_dict name>" (weight norm removed), "ivgan.final_norm.weight/bias" (gpt.final_norm),
 * "ivgan.aa_filter" (12 taps) and "ivgan.upsample_rates" (one float per stage; the state dict does not carry the strides)
 * must be loaded before the build.
 * save_hidden_state (S, gpt_dim) fp32 -- the last two rows are dropped as the graph does -- plus the six per-stage
 * conditioning vectors save_bigvgan_conds_i (C_i floats) and bigvgan_cond_layer_speaker_embedding (C0 floats)
 * -> generated_wav int16 (1, 1, hop*(S-2)+30); *n_out = sample count; wave_host optional as in b200tts_bigvgan_run. */
int b200tts_indextts_vocoder_build(b200tts_engine* e);
int b200tts_indextts_vocoder_run(b200tts_engine* e, const float* hidden_host, int S, const float* const* conds_host,
                                 const float* cond_layer_host, int precision, int16_t* pcm_host, float* wave_host,
                                 int64_t* n_out);

/* ---- IndexTTS GPT-2 acoustic model: graphs B-E (IndexTTS/Export_IndexTTS.py:203-289) and the greedy loop around them
 * (IndexTTS/Inference_IndexTTS_ONNX.py:726-781). The KV cache stays on the device between calls.
 * Tensors "igpt.<name>" must be loaded before the build: text_embedding.weight, text_pos_embedding.emb.weight,
 * mel_embedding.weight, mel_pos_embedding.emb.weight, h.<i>.{ln_1,ln_2}.{weight,bias}, h.<i>.attn.{c_attn,c_proj}.{weight,bias},
 * h.<i>.mlp.{c_fc,c_proj}.{weight,bias} (Hugging Face Conv1D layout (in, out), UNscaled: the build applies the export's
 * head_dim^-0.25 q/k scaling, :250-255), ln_f.*, final_norm.*, mel_head.{weight,bias}, and "igpt.meta" = 8 floats
 * [start_text, stop_text, start_mel, stop_mel, MAX_GENERATE_LENGTH, PENALITY_RANGE, REPEAT_PENALITY, layer-norm eps]. */
int b200tts_indextts_gpt_build(b200tts_engine* e);
/* -> model_dim, layers, heads, mel code count, KV capacity in rows (any pointer may be NULL) */
int b200tts_indextts_gpt_info(b200tts_engine* e, int* dim, int* layers, int* heads, int* mel_codes, int* max_rows);
/* graph B (IndexTTS_B, :203-214): text_ids (n_text) -> text_hidden_state (n_text + 2, dim): start / stop ids are added inside */
int b200tts_indextts_gpt_text_embed(b200tts_engine* e, const int32_t* text_ids_host, int n_text, float* out_host);
/* graph C (IndexTTS_C, :217-225): mel id, gen_len -> gpt_hidden_state (dim); the caller advances gen_len */
int b200tts_indextts_gpt_mel_embed(b200tts_engine* e, int32_t mel_id, int64_t gen_len, float* out_host);
/* graph E (IndexTTS_E.forward, :264-289), one call: hidden_state (ids_len, dim) appended after history_len cached rows
 * (history_len must be the resident length, or 0 to start a sentence); attention_mask = the int8 flag (1 = causal among the
 * new rows); repeat_penality (mel_codes). -> last_hidden_state (dim), max_logit_id, kv_seq_len = history_len + ids_len. */
int b200tts_indextts_gpt_step(b200tts_engine* e, const float* hidden_host, int ids_len, int64_t history_len, int attention_mask,
                              const float* repeat_penality_host, int precision, float* last_hidden_host, int32_t* max_logit_id,
                              int64_t* kv_seq_len);
/* out_key_<layer> (heads, 64, S) and out_value_<layer> (heads, S, 64) of the resident cache, S = resident rows */
int b200tts_indextts_gpt_kv_read(b200tts_engine* e, int layer, float* key_host, float* value_host, int64_t* rows);
/* One sentence with the whole loop on the device (graphs B, C, D, then E until the stop id or the limit
 * min(max_new, MAX_GENERATE_LENGTH - prompt rows); max_new <= 0: no extra limit): conds_latent (cond_rows, dim), text ids
 * -> ids_out (n), hidden_out (n, dim) = last_hidden_state of every call (what the reference concatenates for graph F).
 * repeat_penality_inout (mel_codes, may be NULL = all ones): the reference never resets it between sentences. */
int b200tts_indextts_gpt_generate(b200tts_engine* e, const float* conds_latent_host, int cond_rows, const int32_t* text_ids_host,
                                  int n_text, int max_new, int precision, float* repeat_penality_inout_host, int32_t* ids_out_host,
                                  float* hidden_out_host, int* n_out);
/* same with every buffer on the device (ids_out / hidden_out sized for the limit) */
int b200tts_indextts_gpt_generate_device(b200tts_engine* e, const float* conds_latent_dev, int cond_rows, const int32_t* text_ids_dev,
                                         int n_text, int max_new, int precision, float* repeat_penality_inout_dev,
                                         int32_t* ids_out_dev, float* hidden_out_dev, int* n_out);

/* ---- F5-TTS sessions (F5_TTS/Export_F5.py:98-203, host loop F5-TTS-ONNX-Inference.py:247-311) ----------------
 * Tensors "dit.*" (EMA DiT state dict, Q/K pre-scaled), "vocos.*" (folded) and "f5.*" (export-time constants:
 * time_expand, delta_t, rope_cos/sin, text_pos, stft_basis, fbank, istft_basis, window_sum_inv) must be loaded. */
int b200tts_f5_build(b200tts_engine* e);
/* F5_Preprocess: audio int16 (1,1,L), text_ids int32 (1,n_text), max_duration -> cat_mel_text, cat_mel_text_drop
 * (1,N,612) fp32 and ref_signal_len. noise / rope tables are produced by the caller (session.py) exactly as the
 * graph would: noise is a seeded normal draw, the rope outputs are rows of the constant "f5.rope_*" tables. */
int b200tts_f5_preprocess(b200tts_engine* e, const int16_t* audio_host, int64_t L, const int32_t* text_ids_host,
                          int n_text, int64_t max_duration, float* cat_mel_text_host, float* cat_mel_text_drop_host,
                          int64_t* ref_signal_len);
/* F5_Transformer: n_steps fused NFE steps (FUSE_NFE) from *time_step; noise (1,N,100) updated in place,
 * *time_step += n_steps. rope_cos/rope_sin are the (N,64) rows of rope_cos_q / rope_sin_q [0,0]. */
int b200tts_f5_transformer(b200tts_engine* e, float* noise_host, const float* rope_cos_host, const float* rope_sin_host,
                           const float* cat_mel_text_host, const float* cat_mel_text_drop_host, int N,
                           int32_t* time_step, int n_steps, int precision);
/* F5_Decode: denoised (1,N,100), ref_signal_len -> output_audio int16 (1,1,256*(N-ref-1)); *n_out = sample count.
 * wave_host (optional) receives the pre-cast float (clamped, x32767). */
int b200tts_f5_decode(b200tts_engine* e, const float* denoised_host, int N, int64_t ref_signal_len, int16_t* pcm_host,
                      float* wave_host, int64_t* n_out);
/* Fused fast path = A, then n_steps x B (n_steps < 0: NFE-1), then C with every intermediate resident in HBM.
 * noise_host (N,100) is the Euler start; mel_host (optional, N x 100) receives the denoised mel. */
int b200tts_f5_synthesize(b200tts_engine* e, const int16_t* audio_host, int64_t L, const int32_t* text_ids_host,
                          int n_text, int64_t max_duration, const float* noise_host, int precision, int n_steps,
                          int16_t* pcm_host, int64_t* n_out, float* mel_host);
/* Same with device buffers, enqueued on the engine stream without synchronising (pcm_dev must hold
 * 256*(max_duration - (L/256+1) - 1) samples). */
int b200tts_f5_synthesize_device(b200tts_engine* e, const int16_t* audio_dev, int64_t L, const int32_t* text_ids_dev,
                                 int n_text, int64_t max_duration, const float* noise_dev, int precision, int n_steps,
                                 int16_t* pcm_dev, float* mel_dev);

/* Batched fast path: U utterances that share (L, n_text, max_duration) -- the caller buckets by length, which is how the
 * reference's one-utterance-per-run loop (F5-TTS-ONNX-Inference.py:247-311) is widened for config 4 (many utterances per
 * GPU). Buffers are contiguous per utterance: audio [U][L], text_ids [U][n_text], noise [U][N][100], pcm [U][256*(N-F-1)],
 * mel (optional) [U][N][100]. Results are identical to U calls of b200tts_f5_synthesize_device (the utterances never mix:
 * every kernel is row-wise or per-sequence). bf16 engine only for U > 1. */
int b200tts_f5_synthesize_batch_device(b200tts_engine* e, int U, const int16_t* audio_dev, int64_t L,
                                       const int32_t* text_ids_dev, int n_text, int64_t max_duration, const float* noise_dev,
                                       int precision, int n_steps, int16_t* pcm_dev, float* mel_dev);

/* The metric's pipeline (BASELINE.json: "F5-TTS NFE=32 + BigVGAN 24 kHz"; configs[3]) for U utterances that share
 * (L, n_text, max_duration), in ONE call: graph A per utterance, one batched DiT loop of n_steps Euler steps (< 0: NFE-1), then
 * the GENERATED frames mel[:, ref_len:] (G = max_duration - (L/256+1) of them) through the BigVGAN session
 * (BigVGAN/Export_BigVGAN.py:44-49) without leaving the device. Both b200tts_f5_build and b200tts_bigvgan_build must have run.
 *   audio [U][L] i16, text_ids [U][n_text] i32, noise [U][N][100] f32
 *   -> wav [U][256*G + 30] i16 (BigVGAN), wav_vocos (optional, may be NULL) [U][256*(G-1)] i16 = the reference's own F5_Decode
 *      (Vocos + ISTFT) of the same mel, mel (optional) [U][N][100] f32.
 * The reference never chains these two graphs; parity is per graph: mel vs F5_Transformer, wav vs BigVGAN
 * fed that mel. The _device variant takes device pointers and enqueues on the engine stream without synchronising. */
int b200tts_f5_bigvgan_pipeline(b200tts_engine* e, int U, const int16_t* audio_host, int64_t L, const int32_t* text_ids_host,
                                int n_text, int64_t max_duration, const float* noise_host, int precision, int n_steps,
                                int16_t* wav_host, int16_t* wav_vocos_host, float* mel_host);
int b200tts_f5_bigvgan_pipeline_device(b200tts_engine* e, int U, const int16_t* audio_dev, int64_t L, const int32_t* text_ids_dev,
                                       int n_text, int64_t max_duration, const float* noise_dev, int precision, int n_steps,
                                       int16_t* wav_dev, int16_t* wav_vocos_dev, float* mel_dev);

/* The same pipeline for a RAGGED batch (configs[3]: 64 utterances whose references last 4-8 s): every utterance has its own
 * L[u], n_text[u], max_duration[u] (host arrays of U entries) and all buffers are concatenated in utterance order --
 * audio [sum L], text_ids [sum n_text], noise [sum N][100] -> wav [sum (256*G_u + 30)], wav_vocos (optional) [sum 256*(G_u - 1)],
 * mel (optional) [sum N][100]. The utterances share ONE DiT loop (row-wise GEMMs over all 2*sum(N) rows; attention, RoPE and
 * V^T by per-sequence tables), so every utterance's result equals its single run (reference: one utterance per script run,
 * F5-TTS-ONNX-Inference.py:227-311). */
int b200tts_f5_bigvgan_pipeline_ragged(b200tts_engine* e, int U, const int16_t* audio_host, const int64_t* L,
                                       const int32_t* text_ids_host, const int32_t* n_text, const int64_t* max_duration,
                                       const float* noise_host, int precision, int n_steps, int16_t* wav_host,
                                       int16_t* wav_vocos_host, float* mel_host);
int b200tts_f5_bigvgan_pipeline_ragged_device(b200tts_engine* e, int U, const int16_t* audio_dev, const int64_t* L,
                                              const int32_t* text_ids_dev, const int32_t* n_text, const int64_t* max_duration,
                                              const float* noise_dev, int precision, int n_steps, int16_t* wav_dev,
                                              int16_t* wav_vocos_dev, float* mel_dev);

/* ---- single-op entry points (parity tests of the kernels through the boundary) -------------------------
 * Anti-aliased SnakeBeta (BigVGAN/modeling_modified/act.py:25-29): x (B, C, L) fp32 host in the reference
 * layout -> y (B, C, L) (post=0) or (B, C, L+30) (post=1, the bigvgan.py:370,381-382 tables). alpha_log /
 * beta_log are the raw (logscale) SnakeBeta parameters, taps12 the Kaiser-sinc filter. */
int b200tts_aa_activation(b200tts_engine* e, const float* x_host, int B, int C, int L, const float* alpha_log,
                          const float* beta_log, const float* taps12, int precise, int post, float* y_host);
/* Shifted-row GEMM (csrc/rowgemm.cuh) as a Conv1d on the reference layout: x (B, Cin, L) fp32 host,
 * w (Cout, Cin/groups, k) fp32 host, bias (Cout) or NULL -> y (B, Cout, L) fp32 host; "same" padding
 * (k*dil - dil)/2, stride 1. precision selects the SIMT fp32 or the tcgen05 bf16 kernel. */
int b200tts_conv1d(b200tts_engine* e, const float* x_host, int B, int Cin, int L, const float* w_host, int Cout,
                   int k, int dil, int groups, const float* bias_host, int precision, float* y_host);
/* ConvTranspose1d with kernel 2*stride, padding stride/2 (the BigVGAN upsamplers, bigvgan.py:300-312):
 * x (B, Cin, L), w (Cin, Cout, 2*stride) -> y (B, Cout, stride*L). */
int b200tts_conv_transpose1d(b200tts_engine* e, const float* x_host, int B, int Cin, int L, const float* w_host,
                             int Cout, int stride, const float* bias_host, int precision, float* y_host);

/* The DiT attention kernel alone (F5 modules.py:467: softmax(q @ k, fp32) @ v, no scale, no mask) on the tcgen05
 * path: q, k, v (2, H, N, 64) fp32 host (already roped / pre-scaled) -> out (2, N, H*64) fp32 host. */
int b200tts_attention(b200tts_engine* e, const float* q_host, const float* k_host, const float* v_host, int H, int N,
                      float* out_host);
/* the same with the operand type chosen: precision = B200TTS_BF16 or B200TTS_F16 */
int b200tts_attention_prec(b200tts_engine* e, const float* q_host, const float* k_host, const float* v_host, int H, int N,
                           int precision, float* out_host);

/* Micro-benchmark of the tensor-core shifted-row GEMM alone (tools/bench_gemm.py; not on any product path): B batches of
 * M rows x Cin channels (bf16, synthetic) against taps x N x Cin weights, epilogue = bias (+ fp32 residual/gate when
 * epilogue == 1, bf16 output when epilogue == 2); *ms_out = average CUDA-event milliseconds per launch over iters. */
int b200tts_bench_rowgemm(b200tts_engine* e, int B, int M, int N, int Cin, int taps, int dil, int groups, int epilogue,
                          int iters, float* ms_out);

/* ---- profiling (bench.py roofline leg) ------------------------------------------------------------------
 * Between begin and end every kernel launch is bracketed by CUDA events on the engine stream; end returns a
 * JSON object {"tag": {"launches": n, "ms": total}, ...} valid until the next call on this engine. */
int b200tts_profile_begin(b200tts_engine* e);
const char* b200tts_profile_end(b200tts_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* B200TTS_H */

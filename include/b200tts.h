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

enum { B200TTS_F32 = 0, B200TTS_BF16 = 1 };   /* arithmetic of the dense contractions */

/* ---- lifetime / errors ------------------------------------------------------------------------------ */
/* Replaces InferenceSession construction (F5-TTS-ONNX-Inference.py:173-214): one engine per GPU. */
int b200tts_create(int device, b200tts_engine** out);
void b200tts_destroy(b200tts_engine* e);
const char* b200tts_last_error(void);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the engine's own stream. */
int b200tts_set_stream(b200tts_engine* e, void* cuda_stream);
int b200tts_synchronize(b200tts_engine* e);
/* Number of kernels this library has launched so far in this process (bench.py: gpu_launches). */
unsigned long long b200tts_launch_count(void);

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

/* ---- profiling (bench.py roofline leg) ------------------------------------------------------------------
 * Between begin and end every kernel launch is bracketed by CUDA events on the engine stream; end returns a
 * JSON object {"tag": {"launches": n, "ms": total}, ...} valid until the next call on this engine. */
int b200tts_profile_begin(b200tts_engine* e);
const char* b200tts_profile_end(b200tts_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* B200TTS_H */

// extern "C" boundary of libb200tts (include/b200tts.h). Exceptions never cross it: every entry point
// catches, stores the message thread-locally and returns a non-zero code.
#include "../../include/b200tts.h"

#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>
#include <sstream>

#include "aa_act.cuh"
#include "attention_tc.cuh"
#include "bigvgan.cuh"
#include "dit_chain.cuh"
#include "engine.cuh"
#include "f5.cuh"
#include "gpt2.cuh"
#include "layout.cuh"
#include "rowgemm.cuh"
#include "rowgemm_tc.cuh"

using namespace b200tts;

struct b200tts_engine {
  Engine impl;
};

namespace b200tts {
Engine::~Engine() {
  if (bigvgan) bigvgan_free(bigvgan);
  if (ivgan) bigvgan_free(ivgan);
  if (f5) f5_free(f5);
  if (igpt) gpt_free(igpt);
  for (int i = 0; i < 2; ++i) {
    if (aux_stream[i]) cudaStreamDestroy(aux_stream[i]);
    if (ev_acc[i]) cudaEventDestroy(ev_acc[i]);
    if (ev_end[i]) cudaEventDestroy(ev_end[i]);
  }
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (own_stream && stream) cudaStreamDestroy(stream);
}
void Engine::ensure_aux() {
  if (aux_stream[0]) return;
  for (int i = 0; i < 2; ++i) {
    B2_CUDA(cudaStreamCreateWithFlags(&aux_stream[i], cudaStreamNonBlocking));
    B2_CUDA(cudaEventCreateWithFlags(&ev_acc[i], cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&ev_end[i], cudaEventDisableTiming));
  }
  B2_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
}
}  // namespace b200tts

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& ex) {
    g_last_error = ex.what();
    return 1;
  } catch (...) {
    g_last_error = "unknown error";
    return 2;
  }
}

Engine& eng(b200tts_engine* e) {
  if (!e) fail("null engine handle");
  B2_CUDA(cudaSetDevice(e->impl.device));
  return e->impl;
}

void store_tensor(Engine& E, const char* name, const float* data, const int64_t* shape, int ndim, bool from_device) {
  B2_CHECK(name && data && (shape || ndim == 0), "load_tensor: null argument");
  B2_CHECK(ndim >= 0 && ndim <= 4, "load_tensor: rank must be <= 4");
  Tensor t;
  long n = 1;
  for (int i = 0; i < ndim; ++i) { B2_CHECK(shape[i] > 0, "load_tensor: non-positive dimension"); t.shape.push_back(shape[i]); n *= shape[i]; }
  t.data.alloc((size_t)n);
  B2_CUDA(cudaMemcpyAsync(t.data.p, data, (size_t)n * sizeof(float),
                          from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, E.stream));
  B2_CUDA(cudaStreamSynchronize(E.stream));
  E.weights[name] = std::move(t);
}

}  // namespace

extern "C" {

const char* b200tts_last_error(void) { return g_last_error.c_str(); }

unsigned long long b200tts_launch_count(void) { return g_launch_count; }

int b200tts_debug_chain_plan(int row_blocks, int resident_pairs, int* plan5) {
  if (plan5 == nullptr || row_blocks <= 0 || resident_pairs <= 0) return 1;
  const DitChainPlan pl = dit_chain_plan(row_blocks, resident_pairs);
  plan5[0] = pl.team; plan5[1] = pl.teams; plan5[2] = pl.nrb0; plan5[3] = pl.rem; plan5[4] = pl.team1;
  return 0;
}

int b200tts_create(int device, b200tts_engine** out) {
  return guarded([&] {
    B2_CHECK(out != nullptr, "create: null out pointer");
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0) fail("no CUDA device available (libb200tts has no CPU fallback)");
    B2_CHECK(device >= 0 && device < n, "create: device index out of range");
    B2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) fail(std::string("libb200tts is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
    std::unique_ptr<b200tts_engine> e(new b200tts_engine());
    e->impl.device = device;
    B2_CUDA(cudaStreamCreateWithFlags(&e->impl.stream, cudaStreamNonBlocking));
    e->impl.own_stream = true;
    const char* g = getenv("B200TTS_GRAPHS");           // B200TTS_GRAPHS=0: always enqueue kernel by kernel
    e->impl.graphs.enabled = !(g != nullptr && g[0] == '0');
    const char* ch = getenv("B200TTS_CHAIN");           // B200TTS_CHAIN=0: DiT blocks as separate launches (A/B of dit_chain.cu)
    e->impl.dit_chain = !(ch != nullptr && ch[0] == '0');
    const char* re = getenv("B200TTS_RAGGED_EMBED");    // 0 / 1: per-utterance / one-launch input embedding of a ragged batch
    if (re != nullptr) e->impl.ragged_embed = re[0] != '0';
    *out = e.release();
  });
}

void b200tts_destroy(b200tts_engine* e) {
  if (!e) return;
  cudaSetDevice(e->impl.device);
  cudaDeviceSynchronize();
  delete e;
}

int b200tts_set_stream(b200tts_engine* e, void* cuda_stream) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CUDA(cudaStreamSynchronize(E.stream));
    if (E.own_stream && E.stream) { cudaStreamDestroy(E.stream); E.stream = nullptr; E.own_stream = false; }
    if (cuda_stream) {
      E.stream = (cudaStream_t)cuda_stream;
    } else {
      B2_CUDA(cudaStreamCreateWithFlags(&E.stream, cudaStreamNonBlocking));
      E.own_stream = true;
    }
  });
}

int b200tts_set_option(b200tts_engine* e, const char* name, int value) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(name != nullptr, "set_option: null name");
    B2_CUDA(cudaStreamSynchronize(E.stream));
    const std::string n(name);
    if (n == "dit_chain") E.dit_chain = value != 0;
    else if (n == "cuda_graphs") E.graphs.enabled = value != 0;
    else if (n == "bigvgan_branches") E.bigvgan_branches = value != 0;
    else if (n == "dit_fp8") E.dit_fp8 = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (n == "ragged_embed") E.ragged_embed = value != 0;
    else fail("set_option: unknown option '" + n + "' (known: dit_chain, cuda_graphs, bigvgan_branches, dit_fp8, ragged_embed)");
    E.graphs.clear();                                    // captured graphs bake the code path in
  });
}

int b200tts_synchronize(b200tts_engine* e) {
  return guarded([&] { Engine& E = eng(e); B2_CUDA(cudaStreamSynchronize(E.stream)); });
}

int b200tts_load_tensor(b200tts_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim) {
  return guarded([&] { store_tensor(eng(e), name, host_data, shape, ndim, false); });
}

int b200tts_load_tensor_device(b200tts_engine* e, const char* name, const float* dev_data, const int64_t* shape, int ndim) {
  return guarded([&] { store_tensor(eng(e), name, dev_data, shape, ndim, true); });
}

int b200tts_bigvgan_build(b200tts_engine* e) {
  return guarded([&] {
    Engine& E = eng(e);
    if (E.bigvgan) { bigvgan_free(E.bigvgan); E.bigvgan = nullptr; }
    E.bigvgan = bigvgan_build(E, "bigvgan.");
  });
}

int b200tts_bigvgan_run_device(b200tts_engine* e, const float* mel_dev, int B, int T, int precision, int16_t* pcm_dev,
                               float* wave_dev) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(mel_dev && pcm_dev, "bigvgan_run_device: null buffer");
    B2_CHECK(E.bigvgan != nullptr, "BigVGAN weights are not built (call b200tts_bigvgan_build)");
    run_graphed(E, {10, (long long)(uintptr_t)mel_dev, B, T, precision, (long long)(uintptr_t)pcm_dev, (long long)(uintptr_t)wave_dev},
                [&] { bigvgan_forward(E, *E.bigvgan, mel_dev, B, T, precision, pcm_dev, wave_dev); }, [] {});
  });
}

int b200tts_bigvgan_run(b200tts_engine* e, const float* mel_host, int B, int T, int precision, int16_t* pcm_host,
                        float* wave_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(mel_host && pcm_host, "bigvgan_run: null buffer");
    B2_CHECK(E.bigvgan != nullptr, "BigVGAN weights are not built (call b200tts_bigvgan_build)");
    B2_CHECK(B > 0 && T > 0, "bigvgan_run: empty input");
    const long n_mel = (long)B * T * bigvgan_num_mels(*E.bigvgan);
    const long n_out = (long)B * bigvgan_out_samples(*E.bigvgan, T);
    // persistent staging buffers: stable addresses let repeated calls of one shape replay a captured graph
    E.io_f32a.reserve((size_t)n_mel);
    E.io_i16.reserve((size_t)n_out);
    if (wave_host) E.io_f32b.reserve((size_t)n_out);
    float* d_mel = E.io_f32a.p; int16_t* d_pcm = E.io_i16.p; float* d_wave = wave_host ? E.io_f32b.p : nullptr;
    B2_CUDA(cudaMemcpyAsync(d_mel, mel_host, n_mel * sizeof(float), cudaMemcpyHostToDevice, E.stream));
    run_graphed(E, {11, (long long)(uintptr_t)d_mel, B, T, precision, (long long)(uintptr_t)d_pcm, (long long)(uintptr_t)d_wave},
                [&] { bigvgan_forward(E, *E.bigvgan, d_mel, B, T, precision, d_pcm, d_wave); }, [] {});
    B2_CUDA(cudaMemcpyAsync(pcm_host, d_pcm, n_out * sizeof(int16_t), cudaMemcpyDeviceToHost, E.stream));
    if (wave_host) B2_CUDA(cudaMemcpyAsync(wave_host, d_wave, n_out * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    B2_CUDA(cudaStreamSynchronize(E.stream));
  });
}

int b200tts_indextts_vocoder_build(b200tts_engine* e) {
  return guarded([&] {
    Engine& E = eng(e);
    if (E.ivgan) { bigvgan_free(E.ivgan); E.ivgan = nullptr; }
    B2_CHECK(E.has_weight("ivgan.final_norm.weight"), "ivgan.final_norm.weight (gpt.final_norm) is not loaded");
    E.ivgan = bigvgan_build(E, "ivgan.");
  });
}

int b200tts_indextts_vocoder_run(b200tts_engine* e, const float* hidden_host, int S, const float* const* conds_host,
                                 const float* cond_layer_host, int precision, int16_t* pcm_host, float* wave_host, int64_t* n_out) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.ivgan != nullptr, "IndexTTS vocoder weights are not built (call b200tts_indextts_vocoder_build)");
    B2_CHECK(hidden_host && conds_host && cond_layer_host && pcm_host && n_out, "indextts_vocoder_run: null buffer");
    B2_CHECK(S >= 3, "indextts_vocoder_run: the latent needs at least 3 rows (the last two are dropped)");
    BigVGANModel& M = *E.ivgan;
    cudaStream_t s = E.stream;
    const int T = S - 2, D = bigvgan_num_mels(M), ns = bigvgan_num_stages(M);
    const long n_samples = bigvgan_out_samples(M, T);
    // staging: [latent T*D | cond_0 .. cond_{n-1} | cond_layer] in one fp32 block, PCM in the int16 block
    size_t ncond = (size_t)bigvgan_stage_channels(M, -1);
    for (int i = 0; i < ns; ++i) ncond += (size_t)bigvgan_stage_channels(M, i);
    E.io_f32a.reserve((size_t)T * D + ncond);
    E.io_i16.reserve((size_t)n_samples);
    if (wave_host) E.io_f32b.reserve((size_t)n_samples);
    float* d_lat = E.io_f32a.p;
    std::vector<const float*> d_conds(ns + 1);
    float* cp = d_lat + (size_t)T * D;
    B2_CUDA(cudaMemcpyAsync(d_lat, hidden_host, (size_t)T * D * sizeof(float), cudaMemcpyHostToDevice, s));     // hidden[:-2]
    for (int i = 0; i <= ns; ++i) {
      const int C = bigvgan_stage_channels(M, i < ns ? i : -1);
      const float* src = i < ns ? conds_host[i] : cond_layer_host;
      B2_CHECK(src != nullptr, "indextts_vocoder_run: null conditioning vector");
      B2_CUDA(cudaMemcpyAsync(cp, src, (size_t)C * sizeof(float), cudaMemcpyHostToDevice, s));
      d_conds[i] = cp;
      cp += C;
    }
    int16_t* d_pcm = E.io_i16.p; float* d_wave = wave_host ? E.io_f32b.p : nullptr;
    run_graphed(E, {12, (long long)(uintptr_t)d_lat, T, precision, (long long)(uintptr_t)d_pcm, (long long)(uintptr_t)d_wave},
                [&] { bigvgan_forward(E, M, d_lat, 1, T, precision, d_pcm, d_wave, d_conds.data()); }, [] {});
    B2_CUDA(cudaMemcpyAsync(pcm_host, d_pcm, n_samples * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (wave_host) B2_CUDA(cudaMemcpyAsync(wave_host, d_wave, n_samples * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *n_out = n_samples;
  });
}

int b200tts_indextts_gpt_build(b200tts_engine* e) {
  return guarded([&] {
    Engine& E = eng(e);
    if (E.igpt) { gpt_free(E.igpt); E.igpt = nullptr; }
    E.graphs.clear();
    E.igpt = gpt_build(E);
  });
}

int b200tts_indextts_gpt_info(b200tts_engine* e, int* dim, int* layers, int* heads, int* mel_codes, int* max_rows) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.igpt != nullptr, "IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
    if (dim) *dim = gpt_dim(*E.igpt);
    if (layers) *layers = gpt_layers(*E.igpt);
    if (heads) *heads = gpt_heads(*E.igpt);
    if (mel_codes) *mel_codes = gpt_mel_codes(*E.igpt);
    if (max_rows) *max_rows = gpt_max_rows(*E.igpt);
  });
}

int b200tts_indextts_gpt_text_embed(b200tts_engine* e, const int32_t* text_ids_host, int n_text, float* out_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.igpt != nullptr, "IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
    B2_CHECK(out_host && n_text >= 0 && (text_ids_host || n_text == 0), "gpt_text_embed: null buffer");
    const int D = gpt_dim(*E.igpt);
    cudaStream_t s = E.stream;
    E.io_i32.reserve((size_t)n_text + 1);
    E.io_f32a.reserve((size_t)(n_text + 2) * D);
    if (n_text) B2_CUDA(cudaMemcpyAsync(E.io_i32.p, text_ids_host, (size_t)n_text * sizeof(int), cudaMemcpyHostToDevice, s));
    gpt_text_embed(E, E.io_i32.p, n_text, E.io_f32a.p);
    B2_CUDA(cudaMemcpyAsync(out_host, E.io_f32a.p, (size_t)(n_text + 2) * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_indextts_gpt_mel_embed(b200tts_engine* e, int32_t mel_id, int64_t gen_len, float* out_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.igpt != nullptr, "IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
    B2_CHECK(out_host != nullptr, "gpt_mel_embed: null buffer");
    const int D = gpt_dim(*E.igpt);
    cudaStream_t s = E.stream;
    E.io_i32.reserve(1);
    E.io_f32a.reserve((size_t)D);
    B2_CUDA(cudaMemcpyAsync(E.io_i32.p, &mel_id, sizeof(int), cudaMemcpyHostToDevice, s));
    gpt_mel_embed(E, E.io_i32.p, (int)gen_len, E.io_f32a.p);
    B2_CUDA(cudaMemcpyAsync(out_host, E.io_f32a.p, (size_t)D * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_indextts_gpt_step(b200tts_engine* e, const float* hidden_host, int ids_len, int64_t history_len, int attention_mask,
                              const float* repeat_penality_host, int precision, float* last_hidden_host, int32_t* max_logit_id,
                              int64_t* kv_seq_len) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.igpt != nullptr, "IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
    B2_CHECK(hidden_host && repeat_penality_host && last_hidden_host && max_logit_id && kv_seq_len, "gpt_step: null buffer");
    B2_CHECK(ids_len >= 1 && history_len >= 0 && history_len < (1 << 20), "gpt_step: bad sizes");
    const int D = gpt_dim(*E.igpt), Vm = gpt_mel_codes(*E.igpt);
    cudaStream_t s = E.stream;
    // staging: [hidden rows | penalty | last hidden]
    E.io_f32a.reserve((size_t)ids_len * D + Vm + D);
    E.io_i32.reserve(1);
    float* d_h = E.io_f32a.p; float* d_pen = d_h + (size_t)ids_len * D; float* d_last = d_pen + Vm;
    B2_CUDA(cudaMemcpyAsync(d_h, hidden_host, (size_t)ids_len * D * sizeof(float), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_pen, repeat_penality_host, (size_t)Vm * sizeof(float), cudaMemcpyHostToDevice, s));
    gpt_step(E, d_h, ids_len, (int)history_len, attention_mask, d_pen, precision, d_last, E.io_i32.p);
    B2_CUDA(cudaMemcpyAsync(last_hidden_host, d_last, (size_t)D * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(max_logit_id, E.io_i32.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *kv_seq_len = history_len + ids_len;
  });
}

int b200tts_indextts_gpt_kv_read(b200tts_engine* e, int layer, float* key_host, float* value_host, int64_t* rows) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.igpt != nullptr, "IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
    const int S = gpt_resident_rows(*E.igpt), H = gpt_heads(*E.igpt);
    if (rows) *rows = S;
    if (S == 0 || (!key_host && !value_host)) return;
    B2_CHECK(key_host && value_host, "gpt_kv_read: pass both buffers");
    const size_t n = (size_t)H * S * 64;
    cudaStream_t s = E.stream;
    E.io_f32b.reserve(2 * n);
    gpt_kv_export(E, layer, E.io_f32b.p, E.io_f32b.p + n);
    B2_CUDA(cudaMemcpyAsync(key_host, E.io_f32b.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(value_host, E.io_f32b.p + n, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_indextts_gpt_generate_device(b200tts_engine* e, const float* conds_latent_dev, int cond_rows, const int32_t* text_ids_dev,
                                         int n_text, int max_new, int precision, float* repeat_penality_inout_dev,
                                         int32_t* ids_out_dev, float* hidden_out_dev, int* n_out) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.igpt != nullptr, "IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
    B2_CHECK((conds_latent_dev || cond_rows == 0) && (text_ids_dev || n_text == 0) && ids_out_dev && hidden_out_dev && n_out,
             "gpt_generate: null buffer");
    *n_out = gpt_generate(E, conds_latent_dev, cond_rows, text_ids_dev, n_text, max_new, precision, repeat_penality_inout_dev,
                          ids_out_dev, hidden_out_dev);
    B2_CUDA(cudaStreamSynchronize(E.stream));
  });
}

int b200tts_indextts_gpt_generate(b200tts_engine* e, const float* conds_latent_host, int cond_rows, const int32_t* text_ids_host,
                                  int n_text, int max_new, int precision, float* repeat_penality_inout_host, int32_t* ids_out_host,
                                  float* hidden_out_host, int* n_out) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(E.igpt != nullptr, "IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
    B2_CHECK((conds_latent_host || cond_rows == 0) && (text_ids_host || n_text == 0) && ids_out_host && hidden_out_host && n_out,
             "gpt_generate: null buffer");
    B2_CHECK(cond_rows >= 0 && n_text >= 0, "gpt_generate: bad sizes");
    const int D = gpt_dim(*E.igpt), Vm = gpt_mel_codes(*E.igpt), cap = gpt_max_rows(*E.igpt);
    cudaStream_t s = E.stream;
    // staging: fp32 [conds | penalty | hidden_out (cap + 1 rows)], int32 [text ids | ids_out (cap + 1)]
    E.io_f32a.reserve((size_t)cond_rows * D + Vm + (size_t)(cap + 1) * D);
    E.io_i32.reserve((size_t)n_text + cap + 2);
    float* d_c = E.io_f32a.p; float* d_pen = d_c + (size_t)cond_rows * D; float* d_hid = d_pen + Vm;
    int* d_txt = E.io_i32.p; int* d_ids = d_txt + n_text + 1;
    if (cond_rows) B2_CUDA(cudaMemcpyAsync(d_c, conds_latent_host, (size_t)cond_rows * D * sizeof(float), cudaMemcpyHostToDevice, s));
    if (n_text) B2_CUDA(cudaMemcpyAsync(d_txt, text_ids_host, (size_t)n_text * sizeof(int), cudaMemcpyHostToDevice, s));
    if (repeat_penality_inout_host)
      B2_CUDA(cudaMemcpyAsync(d_pen, repeat_penality_inout_host, (size_t)Vm * sizeof(float), cudaMemcpyHostToDevice, s));
    const int n = gpt_generate(E, d_c, cond_rows, d_txt, n_text, max_new, precision, repeat_penality_inout_host ? d_pen : nullptr,
                               d_ids, d_hid);
    B2_CUDA(cudaMemcpyAsync(ids_out_host, d_ids, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(hidden_out_host, d_hid, (size_t)n * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (repeat_penality_inout_host)
      B2_CUDA(cudaMemcpyAsync(repeat_penality_inout_host, d_pen, (size_t)Vm * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *n_out = n;
  });
}

int b200tts_f5_build(b200tts_engine* e) {
  return guarded([&] {
    Engine& E = eng(e);
    if (E.f5) { f5_free(E.f5); E.f5 = nullptr; }
    E.f5 = f5_build(E);
  });
}

int b200tts_f5_preprocess(b200tts_engine* e, const int16_t* audio_host, int64_t L, const int32_t* text_ids_host, int n_text,
                          int64_t max_duration, float* cat_mel_text_host, float* cat_mel_text_drop_host,
                          int64_t* ref_signal_len) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_host && text_ids_host && cat_mel_text_host && cat_mel_text_drop_host && ref_signal_len, "f5_preprocess: null buffer");
    B2_CHECK(L > 0 && n_text > 0 && max_duration > 0 && max_duration < (1 << 30), "f5_preprocess: bad sizes");
    cudaStream_t s = E.stream;
    E.io_i16.reserve((size_t)L);
    E.io_f32b.reserve((size_t)n_text);
    int16_t* d_audio = E.io_i16.p;
    int* d_ids = reinterpret_cast<int*>(E.io_f32b.p);
    B2_CUDA(cudaMemcpyAsync(d_audio, audio_host, L * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_ids, text_ids_host, n_text * sizeof(int), cudaMemcpyHostToDevice, s));
    f5_preprocess(E, d_audio, L, d_ids, n_text, (int)max_duration);
    const size_t nb = (size_t)max_duration * f5_cond_dim(E) * sizeof(float);
    B2_CUDA(cudaMemcpyAsync(cat_mel_text_host, f5_cond(E), nb, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(cat_mel_text_drop_host, f5_cond_drop(E), nb, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *ref_signal_len = f5_ref_len(E);
  });
}

int b200tts_f5_transformer(b200tts_engine* e, float* noise_host, const float* rope_cos_host, const float* rope_sin_host,
                           const float* cat_mel_text_host, const float* cat_mel_text_drop_host, int N, int32_t* time_step,
                           int n_steps, int precision) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(noise_host && rope_cos_host && rope_sin_host && cat_mel_text_host && cat_mel_text_drop_host && time_step, "f5_transformer: null buffer");
    cudaStream_t s = E.stream;
    f5_begin(E, N);
    const size_t nc = (size_t)N * f5_cond_dim(E) * sizeof(float), nn = (size_t)N * f5_n_mels(E) * sizeof(float);
    float *rc = nullptr, *rs = nullptr;
    f5_rope_buffers(E, &rc, &rs);
    B2_CUDA(cudaMemcpyAsync(f5_noise(E), noise_host, nn, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(f5_cond(E), cat_mel_text_host, nc, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(f5_cond_drop(E), cat_mel_text_drop_host, nc, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(rc, rope_cos_host, (size_t)N * 64 * sizeof(float), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(rs, rope_sin_host, (size_t)N * 64 * sizeof(float), cudaMemcpyHostToDevice, s));
    const int first = *time_step;
    run_graphed(E, {20, N, first, n_steps, precision},
                [&] { f5_prepare_cond(E); f5_steps(E, first, n_steps, precision); }, [] {});
    B2_CUDA(cudaMemcpyAsync(noise_host, f5_noise(E), nn, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *time_step += n_steps;
  });
}

int b200tts_f5_decode(b200tts_engine* e, const float* denoised_host, int N, int64_t ref_signal_len, int16_t* pcm_host,
                      float* wave_host, int64_t* n_out) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(denoised_host && pcm_host && n_out, "f5_decode: null buffer");
    B2_CHECK(N > 0 && ref_signal_len >= 0 && ref_signal_len < N, "f5_decode: ref_signal_len must be in [0, N)");
    cudaStream_t s = E.stream;
    const long ns = 256L * (N - ref_signal_len - 1);
    const size_t nmel = (size_t)N * f5_n_mels(E), nsz = (size_t)(ns > 0 ? ns : 1);
    E.io_f32a.reserve(nmel);
    E.io_i16.reserve(nsz);
    if (wave_host) E.io_f32b.reserve(nsz);
    float* d_mel = E.io_f32a.p; int16_t* d_pcm = E.io_i16.p; float* d_wave = wave_host ? E.io_f32b.p : nullptr;
    B2_CUDA(cudaMemcpyAsync(d_mel, denoised_host, nmel * sizeof(float), cudaMemcpyHostToDevice, s));
    const long got = f5_decode(E, d_mel, N, (int)ref_signal_len, d_pcm, d_wave);
    B2_CHECK(got == ns, "f5_decode: unexpected output length");
    if (ns > 0) {
      B2_CUDA(cudaMemcpyAsync(pcm_host, d_pcm, ns * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
      if (wave_host) B2_CUDA(cudaMemcpyAsync(wave_host, d_wave, ns * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    B2_CUDA(cudaStreamSynchronize(s));
    *n_out = ns;
  });
}

static void synth_device(Engine& E, const int16_t* audio_dev, int64_t L, const int32_t* ids_dev, int n_text, int64_t N,
                         const float* noise_dev, int precision, int n_steps, int16_t* pcm_dev, float* mel_dev) {
  cudaStream_t s = E.stream;
  run_graphed(E, {30, (long long)(uintptr_t)audio_dev, (long long)L, (long long)(uintptr_t)ids_dev, n_text, (long long)N,
                  (long long)(uintptr_t)noise_dev, precision, n_steps, (long long)(uintptr_t)pcm_dev, (long long)(uintptr_t)mel_dev},
              [&] {
                f5_preprocess(E, audio_dev, L, ids_dev, n_text, (int)N, 0, 1, precision);
                B2_CUDA(cudaMemcpyAsync(f5_noise(E), noise_dev, (size_t)N * f5_n_mels(E) * sizeof(float), cudaMemcpyDeviceToDevice, s));
                f5_prepare_cond(E);
                f5_steps(E, 0, n_steps < 0 ? f5_nfe(E) - 1 : n_steps, precision);
                f5_decode(E, nullptr, (int)N, f5_ref_len(E), pcm_dev, nullptr, precision);
                if (mel_dev) B2_CUDA(cudaMemcpyAsync(mel_dev, f5_noise(E), (size_t)N * f5_n_mels(E) * sizeof(float), cudaMemcpyDeviceToDevice, s));
              },
              [&] { f5_restore_shape(E, (int)N, (int)(L / 256 + 1)); });
}

// U utterances that share (L, n_text, max_duration): graph A per utterance, ONE batched DiT loop over all 2U sequences
// (the GEMMs see M = 2*U*N rows instead of 2*N), graph C per utterance.
static void synth_batch_device(Engine& E, int U, const int16_t* audio_dev, int64_t L, const int32_t* ids_dev, int n_text, int64_t N,
                               const float* noise_dev, int precision, int n_steps, int16_t* pcm_dev, float* mel_dev,
                               int16_t* vgan_pcm_dev = nullptr) {
  cudaStream_t s = E.stream;
  const int nm = f5_n_mels(E);
  const int F = (int)(L / 256 + 1);
  const long ns = 256L * (N - F - 1);
  run_graphed(E, {31, U, (long long)(uintptr_t)audio_dev, (long long)L, (long long)(uintptr_t)ids_dev, n_text, (long long)N,
                  (long long)(uintptr_t)noise_dev, precision, n_steps, (long long)(uintptr_t)pcm_dev, (long long)(uintptr_t)mel_dev,
                  (long long)(uintptr_t)vgan_pcm_dev},
              [&] {
                f5_begin(E, (int)N, U);
                for (int u = 0; u < U; ++u)
                  f5_preprocess(E, audio_dev + (size_t)u * L, L, ids_dev + (size_t)u * n_text, n_text, (int)N, u, U, precision);
                B2_CUDA(cudaMemcpyAsync(f5_noise(E), noise_dev, (size_t)U * N * nm * sizeof(float), cudaMemcpyDeviceToDevice, s));
                f5_prepare_cond(E);
                f5_steps(E, 0, n_steps < 0 ? f5_nfe(E) - 1 : n_steps, precision);
                if (pcm_dev)
                  for (int u = 0; u < U; ++u)
                    f5_decode(E, f5_noise(E, u), (int)N, F, pcm_dev + (size_t)u * ns, nullptr, precision);
                if (vgan_pcm_dev)        // the generated frames, channels-last as they lie, straight into BigVGAN (bigvgan.cuh)
                  bigvgan_forward(E, *E.bigvgan, f5_noise(E) + (size_t)F * nm, U, (int)(N - F), precision, vgan_pcm_dev, nullptr, nullptr,
                                  (long)N * nm);
                if (mel_dev) B2_CUDA(cudaMemcpyAsync(mel_dev, f5_noise(E), (size_t)U * N * nm * sizeof(float), cudaMemcpyDeviceToDevice, s));
              },
              [&] { f5_restore_shape(E, (int)N, F, U); });
}

int b200tts_f5_synthesize_batch_device(b200tts_engine* e, int U, const int16_t* audio_dev, int64_t L, const int32_t* text_ids_dev,
                                       int n_text, int64_t max_duration, const float* noise_dev, int precision, int n_steps,
                                       int16_t* pcm_dev, float* mel_dev) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_dev && text_ids_dev && noise_dev && pcm_dev, "f5_synthesize_batch_device: null buffer");
    B2_CHECK(U >= 1 && L > 0 && n_text > 0 && max_duration > L / 256 + 2, "f5_synthesize_batch_device: bad sizes");
    B2_CHECK(precision != PREC_F32 || U == 1, "f5_synthesize_batch_device: the fp32 parity engine takes one utterance at a time");
    synth_batch_device(E, U, audio_dev, L, text_ids_dev, n_text, max_duration, noise_dev, precision, n_steps, pcm_dev, mel_dev);
  });
}

// Ragged batch (BASELINE.json configs[3]: utterances with 4-8 s references share one DiT loop): per-utterance lengths, all
// buffers concatenated in utterance order. pcm_dev (Vocos) and vgan_pcm_dev (BigVGAN) are optional.
static void synth_ragged_device(Engine& E, int U, const int16_t* audio_dev, const int64_t* L, const int32_t* ids_dev, const int32_t* n_text,
                                const int64_t* Ns, const float* noise_dev, int precision, int n_steps, int16_t* pcm_dev, float* mel_dev,
                                int16_t* vgan_pcm_dev) {
  cudaStream_t s = E.stream;
  const int nm = f5_n_mels(E);
  std::vector<int> Nv(U), Fv(U);
  std::vector<long long> key = {32, U, (long long)(uintptr_t)audio_dev, (long long)(uintptr_t)ids_dev, (long long)(uintptr_t)noise_dev, precision,
                                n_steps, (long long)(uintptr_t)pcm_dev, (long long)(uintptr_t)mel_dev, (long long)(uintptr_t)vgan_pcm_dev};
  long Ntot = 0;
  for (int u = 0; u < U; ++u) {
    B2_CHECK(L[u] > 0 && n_text[u] > 0 && Ns[u] > L[u] / 256 + 2 && Ns[u] < (1 << 30), "ragged batch: bad sizes");
    Nv[u] = (int)Ns[u]; Fv[u] = (int)(L[u] / 256 + 1); Ntot += Nv[u];
    key.push_back(L[u]); key.push_back(n_text[u]); key.push_back(Ns[u]);
  }
  f5_begin_ragged(E, U, Nv.data());              // layout + per-row tables: outside the capture (host -> device uploads)
  run_graphed(E, key,
              [&] {
                f5_begin_ragged(E, U, Nv.data());  // (no upload: the tables match)
                size_t ao = 0, io = 0;
                for (int u = 0; u < U; ++u) {
                  f5_preprocess(E, audio_dev + ao, L[u], ids_dev + io, n_text[u], Nv[u], u, U, precision);
                  ao += (size_t)L[u]; io += (size_t)n_text[u];
                }
                B2_CUDA(cudaMemcpyAsync(f5_noise(E), noise_dev, (size_t)Ntot * nm * sizeof(float), cudaMemcpyDeviceToDevice, s));
                f5_prepare_cond(E);
                f5_steps(E, 0, n_steps < 0 ? f5_nfe(E) - 1 : n_steps, precision);
                size_t po = 0, vo = 0;
                for (int u = 0; u < U; ++u) {
                  const int G = Nv[u] - Fv[u];
                  if (pcm_dev) f5_decode(E, f5_noise(E, u), Nv[u], Fv[u], pcm_dev + po, nullptr, precision);
                  if (vgan_pcm_dev)       // generated frames of utterance u, channels-last as they lie
                    bigvgan_forward(E, *E.bigvgan, f5_noise(E, u) + (size_t)Fv[u] * nm, 1, G, precision, vgan_pcm_dev + vo, nullptr, nullptr,
                                    (long)G * nm);
                  po += (size_t)256 * (G - 1);
                  vo += (size_t)bigvgan_out_samples(*E.bigvgan, G);
                }
                if (mel_dev) B2_CUDA(cudaMemcpyAsync(mel_dev, f5_noise(E), (size_t)Ntot * nm * sizeof(float), cudaMemcpyDeviceToDevice, s));
              },
              [&] { f5_restore_ragged(E, U, Nv.data(), Fv.data()); });
}

static void pipeline_check(Engine& E, int U, int64_t L, int n_text, int64_t N, int precision) {
  B2_CHECK(E.f5 != nullptr, "F5 weights are not built (call b200tts_f5_build)");
  B2_CHECK(E.bigvgan != nullptr, "BigVGAN weights are not built (call b200tts_bigvgan_build)");
  B2_CHECK(U >= 1 && L > 0 && n_text > 0 && N > L / 256 + 2 && N < (1 << 30), "f5_bigvgan_pipeline: bad sizes");
  B2_CHECK(precision != PREC_F32 || U == 1, "f5_bigvgan_pipeline: the fp32 parity engine takes one utterance at a time");
  B2_CHECK(bigvgan_num_mels(*E.bigvgan) == f5_n_mels(E), "f5_bigvgan_pipeline: the vocoder's mel width differs from the DiT's");
}

int b200tts_f5_bigvgan_pipeline_device(b200tts_engine* e, int U, const int16_t* audio_dev, int64_t L, const int32_t* text_ids_dev,
                                       int n_text, int64_t max_duration, const float* noise_dev, int precision, int n_steps,
                                       int16_t* wav_dev, int16_t* wav_vocos_dev, float* mel_dev) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_dev && text_ids_dev && noise_dev && wav_dev, "f5_bigvgan_pipeline_device: null buffer");
    pipeline_check(E, U, L, n_text, max_duration, precision);
    synth_batch_device(E, U, audio_dev, L, text_ids_dev, n_text, max_duration, noise_dev, precision, n_steps, wav_vocos_dev, mel_dev, wav_dev);
  });
}

int b200tts_f5_bigvgan_pipeline(b200tts_engine* e, int U, const int16_t* audio_host, int64_t L, const int32_t* text_ids_host, int n_text,
                                int64_t max_duration, const float* noise_host, int precision, int n_steps, int16_t* wav_host,
                                int16_t* wav_vocos_host, float* mel_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_host && text_ids_host && noise_host && wav_host, "f5_bigvgan_pipeline: null buffer");
    pipeline_check(E, U, L, n_text, max_duration, precision);
    cudaStream_t s = E.stream;
    const int64_t N = max_duration, F = L / 256 + 1, G = N - F;
    const int nm = f5_n_mels(E);
    const long nv = bigvgan_out_samples(*E.bigvgan, (int)G), ns = 256L * (G - 1);
    // persistent staging (stable addresses -> graph replay): i16 [audio | wav | wav_vocos], f32 [noise], f32 [mel | ids]
    const size_t Lp = (size_t)round_up((long)U * L, 8), nvp = (size_t)round_up((long)U * nv, 8), nsp = (size_t)round_up((long)U * ns, 8);
    const size_t nmel = (size_t)U * N * nm;
    E.io_i16.reserve(Lp + nvp + nsp);
    E.io_f32a.reserve(nmel);
    E.io_f32b.reserve(nmel + (size_t)round_up((long)U * n_text, 4));
    int16_t* d_audio = E.io_i16.p; int16_t* d_wav = d_audio + Lp; int16_t* d_voc = wav_vocos_host ? d_wav + nvp : nullptr;
    float* d_noise = E.io_f32a.p; float* d_mel = mel_host ? E.io_f32b.p : nullptr;
    int* d_ids = reinterpret_cast<int*>(E.io_f32b.p + nmel);
    B2_CUDA(cudaMemcpyAsync(d_audio, audio_host, (size_t)U * L * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_ids, text_ids_host, (size_t)U * n_text * sizeof(int), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_noise, noise_host, nmel * sizeof(float), cudaMemcpyHostToDevice, s));
    synth_batch_device(E, U, d_audio, L, d_ids, n_text, N, d_noise, precision, n_steps, d_voc, d_mel, d_wav);
    B2_CUDA(cudaMemcpyAsync(wav_host, d_wav, (size_t)U * nv * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (wav_vocos_host) B2_CUDA(cudaMemcpyAsync(wav_vocos_host, d_voc, (size_t)U * ns * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (mel_host) B2_CUDA(cudaMemcpyAsync(mel_host, d_mel, nmel * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_f5_bigvgan_pipeline_ragged_device(b200tts_engine* e, int U, const int16_t* audio_dev, const int64_t* L,
                                              const int32_t* text_ids_dev, const int32_t* n_text, const int64_t* max_duration,
                                              const float* noise_dev, int precision, int n_steps, int16_t* wav_dev, int16_t* wav_vocos_dev,
                                              float* mel_dev) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_dev && L && text_ids_dev && n_text && max_duration && noise_dev && wav_dev, "f5_bigvgan_pipeline_ragged_device: null buffer");
    B2_CHECK(U >= 1 && U <= 4096, "f5_bigvgan_pipeline_ragged_device: batch size");
    B2_CHECK(E.f5 != nullptr && E.bigvgan != nullptr, "F5 / BigVGAN weights are not built");
    B2_CHECK(precision != PREC_F32 || U == 1, "the fp32 parity engine takes one utterance at a time");
    synth_ragged_device(E, U, audio_dev, L, text_ids_dev, n_text, max_duration, noise_dev, precision, n_steps, wav_vocos_dev, mel_dev, wav_dev);
  });
}

int b200tts_f5_bigvgan_pipeline_ragged(b200tts_engine* e, int U, const int16_t* audio_host, const int64_t* L, const int32_t* text_ids_host,
                                       const int32_t* n_text, const int64_t* max_duration, const float* noise_host, int precision, int n_steps,
                                       int16_t* wav_host, int16_t* wav_vocos_host, float* mel_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_host && L && text_ids_host && n_text && max_duration && noise_host && wav_host, "f5_bigvgan_pipeline_ragged: null buffer");
    B2_CHECK(U >= 1 && U <= 4096, "f5_bigvgan_pipeline_ragged: batch size");
    B2_CHECK(E.f5 != nullptr && E.bigvgan != nullptr, "F5 / BigVGAN weights are not built");
    B2_CHECK(precision != PREC_F32 || U == 1, "the fp32 parity engine takes one utterance at a time");
    cudaStream_t s = E.stream;
    const int nm = f5_n_mels(E);
    long Ltot = 0, ttot = 0, Ntot = 0, nv = 0, ns = 0;
    for (int u = 0; u < U; ++u) {
      B2_CHECK(L[u] > 0 && n_text[u] > 0 && max_duration[u] > L[u] / 256 + 2, "f5_bigvgan_pipeline_ragged: bad sizes");
      const long G = max_duration[u] - (L[u] / 256 + 1);
      Ltot += L[u]; ttot += n_text[u]; Ntot += max_duration[u];
      nv += bigvgan_out_samples(*E.bigvgan, (int)G); ns += 256L * (G - 1);
    }
    // persistent staging (stable addresses -> graph replay): i16 [audio | wav | wav_vocos], f32 [noise], f32 [mel | ids]
    const size_t Lp = (size_t)round_up(Ltot, 8), nvp = (size_t)round_up(nv, 8), nsp = (size_t)round_up(ns, 8), nmel = (size_t)Ntot * nm;
    E.io_i16.reserve(Lp + nvp + nsp);
    E.io_f32a.reserve(nmel);
    E.io_f32b.reserve(nmel + (size_t)round_up(ttot, 4));
    int16_t* d_audio = E.io_i16.p; int16_t* d_wav = d_audio + Lp; int16_t* d_voc = wav_vocos_host ? d_wav + nvp : nullptr;
    float* d_noise = E.io_f32a.p; float* d_mel = mel_host ? E.io_f32b.p : nullptr;
    int* d_ids = reinterpret_cast<int*>(E.io_f32b.p + nmel);
    B2_CUDA(cudaMemcpyAsync(d_audio, audio_host, (size_t)Ltot * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_ids, text_ids_host, (size_t)ttot * sizeof(int), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_noise, noise_host, nmel * sizeof(float), cudaMemcpyHostToDevice, s));
    synth_ragged_device(E, U, d_audio, L, d_ids, n_text, max_duration, d_noise, precision, n_steps, d_voc, d_mel, d_wav);
    B2_CUDA(cudaMemcpyAsync(wav_host, d_wav, (size_t)nv * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (wav_vocos_host) B2_CUDA(cudaMemcpyAsync(wav_vocos_host, d_voc, (size_t)ns * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (mel_host) B2_CUDA(cudaMemcpyAsync(mel_host, d_mel, nmel * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_f5_synthesize_device(b200tts_engine* e, const int16_t* audio_dev, int64_t L, const int32_t* text_ids_dev,
                                 int n_text, int64_t max_duration, const float* noise_dev, int precision, int n_steps,
                                 int16_t* pcm_dev, float* mel_dev) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_dev && text_ids_dev && noise_dev && pcm_dev, "f5_synthesize_device: null buffer");
    synth_device(E, audio_dev, L, text_ids_dev, n_text, max_duration, noise_dev, precision, n_steps, pcm_dev, mel_dev);
  });
}

int b200tts_f5_synthesize(b200tts_engine* e, const int16_t* audio_host, int64_t L, const int32_t* text_ids_host, int n_text,
                          int64_t max_duration, const float* noise_host, int precision, int n_steps, int16_t* pcm_host,
                          int64_t* n_out, float* mel_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(audio_host && text_ids_host && noise_host && pcm_host && n_out, "f5_synthesize: null buffer");
    B2_CHECK(L > 0 && n_text > 0 && max_duration > 0 && max_duration < (1 << 30), "f5_synthesize: bad sizes");
    cudaStream_t s = E.stream;
    const int64_t N = max_duration, F = L / 256 + 1;
    const long ns = 256L * (N - F - 1);
    B2_CHECK(ns > 0, "f5_synthesize: max_duration leaves no frames to generate");
    const int nm = f5_n_mels(E);
    // persistent staging (stable addresses -> graph replay): [audio i16 | pcm i16] and [noise | mel | ids] blocks
    const size_t Lp = (size_t)round_up(L, 8), nsp = (size_t)round_up(ns, 8), nmel = (size_t)N * nm;
    E.io_i16.reserve(Lp + nsp);
    E.io_f32a.reserve(nmel);
    E.io_f32b.reserve(nmel + (size_t)round_up(n_text, 4));
    int16_t* d_audio = E.io_i16.p; int16_t* d_pcm = E.io_i16.p + Lp;
    float* d_noise = E.io_f32a.p; float* d_mel = mel_host ? E.io_f32b.p : nullptr;
    int* d_ids = reinterpret_cast<int*>(E.io_f32b.p + nmel);
    B2_CUDA(cudaMemcpyAsync(d_audio, audio_host, L * sizeof(int16_t), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_ids, text_ids_host, n_text * sizeof(int), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_noise, noise_host, nmel * sizeof(float), cudaMemcpyHostToDevice, s));
    synth_device(E, d_audio, L, d_ids, n_text, N, d_noise, precision, n_steps, d_pcm, d_mel);
    B2_CUDA(cudaMemcpyAsync(pcm_host, d_pcm, ns * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (mel_host) B2_CUDA(cudaMemcpyAsync(mel_host, d_mel, nmel * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *n_out = ns;
  });
}

int b200tts_aa_activation(b200tts_engine* e, const float* x_host, int B, int C, int L, const float* alpha_log,
                          const float* beta_log, const float* taps12, int precise, int post, float* y_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(x_host && alpha_log && beta_log && taps12 && y_host, "aa_activation: null buffer");
    B2_CHECK(B > 0 && C > 0 && L > 0, "aa_activation: empty input");
    cudaStream_t s = E.stream;
    const int Lo = post ? L + 30 : L;
    const int Cp = (int)round_up(C, 2);          // the kernel walks channel pairs: an odd C gets one zero channel
    DevBuf<float> x((size_t)B * Cp * L), xt((size_t)B * Cp * L), y((size_t)B * Cp * Lo), yt((size_t)B * Cp * Lo), al(Cp), ib(Cp);
    std::vector<float> ha(Cp, 1.0f), hb(Cp, 1.0f);
    for (int i = 0; i < C; ++i) { ha[i] = expf(alpha_log[i]); hb[i] = 1.0f / (expf(beta_log[i]) + 1e-9f); }
    if (Cp == C) {
      B2_CUDA(cudaMemcpyAsync(x.p, x_host, x.n * sizeof(float), cudaMemcpyHostToDevice, s));
    } else {
      B2_CUDA(cudaMemsetAsync(x.p, 0, x.n * sizeof(float), s));
      B2_CUDA(cudaMemcpy2DAsync(x.p, (size_t)Cp * L * sizeof(float), x_host, (size_t)C * L * sizeof(float), (size_t)C * L * sizeof(float),
                                B, cudaMemcpyHostToDevice, s));
    }
    B2_CUDA(cudaMemcpyAsync(al.p, ha.data(), Cp * sizeof(float), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(ib.p, hb.data(), Cp * sizeof(float), cudaMemcpyHostToDevice, s));
    aa_set_filter(taps12);
    batched_transpose(x.p, xt.p, B, Cp, L, s);                      // (B,C,L) -> (B,L,C)
    aa_snake(xt.p, 0, yt.p, 0, al.p, ib.p, B, Cp, L, precise != 0, post != 0, s);
    batched_transpose(yt.p, y.p, B, Lo, Cp, s);                     // (B,Lo,C) -> (B,C,Lo)
    B2_CUDA(cudaMemcpy2DAsync(y_host, (size_t)C * Lo * sizeof(float), y.p, (size_t)Cp * Lo * sizeof(float), (size_t)C * Lo * sizeof(float),
                              B, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_conv1d(b200tts_engine* e, const float* x_host, int B, int Cin, int L, const float* w_host, int Cout, int k,
                   int dil, int groups, const float* bias_host, int precision, float* y_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(x_host && w_host && y_host, "conv1d: null buffer");
    B2_CHECK(B > 0 && Cin > 0 && L > 0 && Cout > 0 && k > 0 && dil > 0 && groups > 0, "conv1d: bad shape");
    B2_CHECK(Cin % groups == 0 && Cout % groups == 0, "conv1d: channels must divide by groups");
    B2_CHECK(k % 2 == 1, "conv1d: 'same' padding needs an odd kernel");
    cudaStream_t s = E.stream;
    const int cg = Cin / groups, ng = Cout / groups;
    // weights (Cout, cg, k) -> [g][j][c][n]
    std::vector<float> wf((size_t)groups * k * cg * ng), wt((size_t)groups * k * ng * cg);
    for (int g = 0; g < groups; ++g)
      for (int n = 0; n < ng; ++n)
        for (int c = 0; c < cg; ++c)
          for (int j = 0; j < k; ++j) {
            const float v = w_host[((size_t)(g * ng + n) * cg + c) * k + j];
            wf[(((size_t)g * k + j) * cg + c) * ng + n] = v;
            wt[(((size_t)g * k + j) * ng + n) * cg + c] = v;
          }
    DevBuf<float> x((size_t)B * Cin * L), xt((size_t)B * Cin * L), y((size_t)B * Cout * L), yt((size_t)B * Cout * L), w(wf.size()), bias;
    B2_CUDA(cudaMemcpyAsync(x.p, x_host, x.n * sizeof(float), cudaMemcpyHostToDevice, s));
    if (bias_host) { bias.alloc(Cout); B2_CUDA(cudaMemcpyAsync(bias.p, bias_host, Cout * sizeof(float), cudaMemcpyHostToDevice, s)); }
    batched_transpose(x.p, xt.p, B, Cin, L, s);
    RowGemm p;
    p.x_bstride = (long)L * Cin; p.ldx = Cin; p.Lin = L;
    p.Cin = cg; p.N = ng; p.taps = k; p.dil = dil; p.center = (k - 1) / 2; p.groups = groups;
    p.M = L; p.B = B;
    p.out = yt.p; p.o_bstride = (long)L * Cout; p.ldo = Cout;
    p.bias = bias_host ? bias.p : nullptr;
    if (precision == PREC_F32) {
      B2_CUDA(cudaMemcpyAsync(w.p, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice, s));
      p.x = xt.p; p.w = w.p; p.ldw = ng;
      rowgemm_f32(p, s);
    } else {
      B2_CUDA(cudaMemcpyAsync(w.p, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice, s));
      const int f16 = precision == PREC_F16 ? 1 : 0;
      TcWeight tw;
      tc_weight_from_f32(tw, w.p, groups, k, ng, cg, s, f16);
      const int ld = (int)round_up(Cin, 8);
      DevBuf<__nv_bfloat16> x16((size_t)B * L * ld);
      cast_pad_f32_to_bf16(xt.p, x16.p, (long)B * L, Cin, ld, s, f16);
      p.x = x16.p; p.ldx = ld; p.x_bstride = (long)L * ld; p.f16 = f16;
      rowgemm_tc(p, tw, s);
      B2_CUDA(cudaStreamSynchronize(s));   // tw / x16 are freed at scope exit
    }
    batched_transpose(yt.p, y.p, B, L, Cout, s);
    B2_CUDA(cudaMemcpyAsync(y_host, y.p, y.n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_conv_transpose1d(b200tts_engine* e, const float* x_host, int B, int Cin, int L, const float* w_host,
                             int Cout, int stride, const float* bias_host, int precision, float* y_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(x_host && w_host && y_host, "conv_transpose1d: null buffer");
    B2_CHECK(B > 0 && Cin > 0 && L > 0 && Cout > 0 && stride > 0 && stride % 2 == 0, "conv_transpose1d: bad shape");
    cudaStream_t s = E.stream;
    const int u = stride;
    const long N = (long)u * Cout, Lo = (long)L * u;
    // w (Cin, Cout, 2u) -> [tap][c][r*Cout+n] and its [tap][r*Cout+n][c] transpose
    std::vector<float> wf(2 * (size_t)Cin * N), wt(2 * (size_t)Cin * N), bf((size_t)N, 0.f);
    for (int tap = 0; tap < 2; ++tap)
      for (int c = 0; c < Cin; ++c)
        for (int r = 0; r < u; ++r)
          for (int n = 0; n < Cout; ++n) {
            const float v = w_host[((size_t)c * Cout + n) * (2 * u) + (tap == 0 ? r + u : r)];
            wf[((size_t)tap * Cin + c) * N + (size_t)r * Cout + n] = v;
            wt[((size_t)tap * N + (size_t)r * Cout + n) * Cin + c] = v;
          }
    if (bias_host)
      for (long i = 0; i < N; ++i) bf[i] = bias_host[i % Cout];
    DevBuf<float> x((size_t)B * Cin * L), xt((size_t)B * Cin * L), y((size_t)B * Cout * Lo), yt((size_t)B * Cout * Lo), w(wf.size()), bias((size_t)N);
    B2_CUDA(cudaMemcpyAsync(x.p, x_host, x.n * sizeof(float), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(bias.p, bf.data(), N * sizeof(float), cudaMemcpyHostToDevice, s));
    batched_transpose(x.p, xt.p, B, Cin, L, s);
    RowGemm p;
    p.x_bstride = (long)L * Cin; p.ldx = Cin; p.Lin = L;
    p.Cin = Cin; p.N = (int)N; p.taps = 2; p.dil = 1; p.center = 1;
    p.M = L + 1; p.B = B;
    p.out = yt.p; p.o_bstride = Lo * Cout; p.ldo = (int)N;
    p.o_shift = -(long)(u / 2) * Cout; p.o_limit = Lo * Cout;
    p.bias = bias.p;
    if (precision == PREC_F32) {
      B2_CUDA(cudaMemcpyAsync(w.p, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice, s));
      p.x = xt.p; p.w = w.p; p.ldw = (int)N;
      rowgemm_f32(p, s);
    } else {
      B2_CUDA(cudaMemcpyAsync(w.p, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice, s));
      const int f16 = precision == PREC_F16 ? 1 : 0;
      TcWeight tw;
      tc_weight_from_f32(tw, w.p, 1, 2, (int)N, Cin, s, f16);
      const int ld = (int)round_up(Cin, 8);
      DevBuf<__nv_bfloat16> x16((size_t)B * L * ld);
      cast_pad_f32_to_bf16(xt.p, x16.p, (long)B * L, Cin, ld, s, f16);
      p.x = x16.p; p.ldx = ld; p.x_bstride = (long)L * ld; p.f16 = f16;
      rowgemm_tc(p, tw, s);
      B2_CUDA(cudaStreamSynchronize(s));
    }
    batched_transpose(yt.p, y.p, B, (int)Lo, Cout, s);
    B2_CUDA(cudaMemcpyAsync(y_host, y.p, y.n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_attention(b200tts_engine* e, const float* q_host, const float* k_host, const float* v_host, int H, int N,
                      float* out_host) {
  return b200tts_attention_prec(e, q_host, k_host, v_host, H, N, B200TTS_BF16, out_host);
}

int b200tts_attention_prec(b200tts_engine* e, const float* q_host, const float* k_host, const float* v_host, int H, int N,
                           int precision, float* out_host) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(q_host && k_host && v_host && out_host, "attention: null buffer");
    B2_CHECK(precision == PREC_BF16 || precision == PREC_F16, "attention: the tcgen05 kernel takes bf16 or fp16 operands");
    const int f16 = precision == PREC_F16 ? 1 : 0;
    B2_CHECK(H > 0 && N > 0, "attention: empty problem");
    cudaStream_t s = E.stream;
    const int D = H * 64, Np = (int)round_up(N, 8);
    // host-side relayout into the engine's operand formats: qk [2][N][2D] (q | k), vT [2H][64][Np]
    std::vector<float> qk((size_t)2 * N * 2 * D), vt((size_t)2 * H * 64 * Np, 0.f);
    for (int b = 0; b < 2; ++b)
      for (int h = 0; h < H; ++h)
        for (int t = 0; t < N; ++t)
          for (int d = 0; d < 64; ++d) {
            const size_t src = (((size_t)b * H + h) * N + t) * 64 + d;
            qk[((size_t)b * N + t) * 2 * D + h * 64 + d] = q_host[src];
            qk[((size_t)b * N + t) * 2 * D + D + h * 64 + d] = k_host[src];
            vt[(((size_t)b * H + h) * 64 + d) * Np + t] = v_host[src];
          }
    DevBuf<float> d_qk(qk.size()), d_vt(vt.size()), d_o32((size_t)2 * N * D);
    DevBuf<__nv_bfloat16> qk16(qk.size()), vt16(vt.size()), o16((size_t)2 * N * D);
    B2_CUDA(cudaMemcpyAsync(d_qk.p, qk.data(), qk.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d_vt.p, vt.data(), vt.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    cast_f32_to_bf16(d_qk.p, qk16.p, (long)qk.size(), s, f16);
    cast_f32_to_bf16(d_vt.p, vt16.p, (long)vt.size(), s, f16);
    attention_tc(qk16.p, vt16.p, Np, o16.p, 2, N, H, s, f16);
    cast_bf16_to_f32(o16.p, d_o32.p, (long)2 * N * D, s, f16);
    B2_CUDA(cudaMemcpyAsync(out_host, d_o32.p, d_o32.n * sizeof(float), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
  });
}

int b200tts_bench_rowgemm(b200tts_engine* e, int B, int M, int N, int Cin, int taps, int dil, int groups, int epilogue,
                          int iters, float* ms_out) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CHECK(ms_out && B > 0 && M > 0 && N > 0 && Cin > 0 && taps > 0 && dil > 0 && groups > 0 && iters > 0, "bench_rowgemm: bad arguments");
    cudaStream_t s = E.stream;
    const int ldx = (int)round_up((long)groups * Cin, 8);
    const size_t xn = (size_t)B * M * ldx, wn = (size_t)groups * taps * N * Cin, on = (size_t)B * M * groups * N;
    DevBuf<float> xf(xn), wf(wn), bias((size_t)groups * N), outf(on), resf(on);
    DevBuf<__nv_bfloat16> x16(xn), out16(on);
    std::vector<float> h(std::max(xn, wn));
    unsigned st = 12345u;
    auto rnd = [&] { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (size_t i = 0; i < xn; ++i) h[i] = rnd();
    B2_CUDA(cudaMemcpyAsync(xf.p, h.data(), xn * sizeof(float), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < wn; ++i) h[i] = rnd() * 0.05f;
    B2_CUDA(cudaMemcpyAsync(wf.p, h.data(), wn * sizeof(float), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemsetAsync(bias.p, 0, bias.n * sizeof(float), s));
    B2_CUDA(cudaMemsetAsync(resf.p, 0, on * sizeof(float), s));
    cast_f32_to_bf16(xf.p, x16.p, (long)xn, s);
    TcWeight tw;
    tc_weight_from_f32(tw, wf.p, groups, taps, N, Cin, s);
    RowGemm p;
    p.x = x16.p; p.x_bstride = (long)M * ldx; p.ldx = ldx; p.Lin = M;
    p.Cin = Cin; p.N = N; p.taps = taps; p.dil = dil; p.center = (taps - 1) / 2; p.groups = groups; p.M = M; p.B = B;
    // epilogue: 0 bias, fp32 out | 1 + fp32 residual + gate | 2 bf16 out | 3 fp32 out + residual from a separate buffer |
    //           4 bf16 out + fp32 residual
    const bool o16 = epilogue == 2 || epilogue == 4;
    p.out = o16 ? (void*)out16.p : (void*)outf.p; p.out_bf16 = o16;
    p.o_bstride = (long)M * groups * N; p.ldo = groups * N; p.bias = bias.p;
    if (epilogue == 1) { p.res = outf.p; p.gate = bias.p; }
    if (epilogue == 3 || epilogue == 4) p.res = resf.p;
    for (int i = 0; i < 3; ++i) rowgemm_tc(p, tw, s);
    cudaEvent_t a, b;
    B2_CUDA(cudaEventCreate(&a)); B2_CUDA(cudaEventCreate(&b));
    B2_CUDA(cudaEventRecord(a, s));
    for (int i = 0; i < iters; ++i) rowgemm_tc(p, tw, s);
    B2_CUDA(cudaEventRecord(b, s));
    B2_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    B2_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_out = ms / iters;
  });
}

int b200tts_profile_begin(b200tts_engine* e) {
  return guarded([&] {
    Engine& E = eng(e);
    B2_CUDA(cudaStreamSynchronize(E.stream));
    E.prof.collect();
    E.prof.enabled = true;
  });
}

const char* b200tts_profile_end(b200tts_engine* e) {
  if (!e) return "{}";
  int rc = guarded([&] {
    Engine& E = eng(e);
    E.prof.enabled = false;
    auto m = E.prof.collect();
    std::ostringstream os;
    os << "{";
    bool first = true;
    for (auto& kv : m) {
      if (!first) os << ", ";
      first = false;
      os << "\"" << kv.first << "\": {\"launches\": " << kv.second.first << ", \"ms\": " << kv.second.second << "}";
    }
    os << "}";
    E.prof_report = os.str();
  });
  return rc == 0 ? e->impl.prof_report.c_str() : "{}";
}

}  // extern "C"

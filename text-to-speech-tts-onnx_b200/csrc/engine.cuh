// Engine state behind the C ABI (include/b200tts.h): one engine per GPU, one stream, named weight
// store, grow-only workspaces, optional per-kernel CUDA-event profiling.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

namespace b200tts {

struct Tensor {
  std::vector<long> shape;
  DevBuf<float> data;
  long numel() const { long n = 1; for (long s : shape) n *= s; return n; }
};

// Per-tag CUDA-event timing of kernel launches on the engine stream (bench.py roofline leg).
struct Profiler {
  bool enabled = false;
  struct Rec { std::string tag; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  // -> index of the record (scopes may nest: each one closes its own record), -1 when disabled
  int begin(const char* tag, cudaStream_t s) {
    if (!enabled) return -1;
    Rec r; r.tag = tag;
    B2_CUDA(cudaEventCreate(&r.a)); B2_CUDA(cudaEventCreate(&r.b));
    B2_CUDA(cudaEventRecord(r.a, s));
    recs.push_back(r);
    return (int)recs.size() - 1;
  }
  void end(int idx, cudaStream_t s) {
    if (idx < 0 || idx >= (int)recs.size()) return;
    cudaEventRecord(recs[idx].b, s);
  }
  // -> map tag -> (count, total ms); destroys the events
  std::map<std::string, std::pair<long, double>> collect() {
    std::map<std::string, std::pair<long, double>> out;
    for (auto& r : recs) {
      B2_CUDA(cudaEventSynchronize(r.b));
      float ms = 0.f;
      B2_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      auto& e = out[r.tag];
      e.first += 1; e.second += ms;
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    recs.clear();
    return out;
  }
};

struct ProfScope {
  Profiler& p; cudaStream_t s; int idx;
  ProfScope(Profiler& p_, const char* tag, cudaStream_t s_) : p(p_), s(s_), idx(p_.begin(tag, s_)) {}
  ~ProfScope() { p.end(idx, s); }
};

// CUDA-graph cache: the hot path is thousands of short kernels per utterance (5138 for one F5 utterance), and enqueueing
// them costs more host time (24 us / launch measured) than the GPU needs to run them. A call is identified by a key
// (entry point, shapes, buffer addresses); the first call with a key runs eagerly (and grows workspaces / builds lazy weight
// layouts), the second captures the same enqueue sequence into a graph, later ones replay it with one launch.
struct GraphCache {
  struct Entry {
    std::vector<long long> key;
    unsigned long long epoch = 0;      // g_alloc_epoch after the last eager run / at capture
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches = 0;   // kernels inside the graph (for b200tts_launch_count)
  };
  std::vector<Entry> entries;
  bool enabled = true;
  Entry* find(const std::vector<long long>& key) {
    for (auto& e : entries) if (e.key == key) return &e;
    return nullptr;
  }
  void clear() {
    for (auto& e : entries) if (e.exec) cudaGraphExecDestroy(e.exec);
    entries.clear();
  }
  ~GraphCache() { clear(); }
};

struct BigVGANModel;
struct F5Model;
struct GptModel;

struct Engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::map<std::string, Tensor> weights;   // reference state_dict names, prefixed "bigvgan." / "dit." / "vocos."
  Profiler prof;
  GraphCache graphs;
  DevBuf<float> io_f32a, io_f32b;          // persistent staging of the host-pointer entry points (stable addresses for graphs)
  DevBuf<int16_t> io_i16;
  DevBuf<int> io_i32;
  std::string prof_report;                 // last JSON report (owned here so the C ABI can hand out a pointer)
  BigVGANModel* bigvgan = nullptr;         // owned; freed by bigvgan_free / f5_free in ~Engine (api.cu)
  BigVGANModel* ivgan = nullptr;           // the IndexTTS_F vocoder (same generator family, see bigvgan.cuh)
  F5Model* f5 = nullptr;
  GptModel* igpt = nullptr;                // IndexTTS GPT-2 acoustic model (gpt2.cuh)
  bool dit_chain = true;                   // F5 DiT blocks through the fused row-block chain kernel (dit_chain.cu); b200tts_set_option
  // ff1 and q|k|v of the fused chain with e4m3 operands (tcgen05 kind::f8f6f4, fp32 accumulation): an OPTIONAL lower-fidelity
  // mode (PCM SNR ~30 dB instead of ~60, oracle/fp8_study.py); off by default. b200tts_set_option("dit_fp8", 1)
  int dit_fp8 = 0;                         // 0 off, 1 ff1 + q|k|v, 2 ff2 as well
  // BigVGAN: the three resblocks of a stage (kernel sizes 3 / 7 / 11, bigvgan.py:396-399) are independent until their sum, so
  // they run as three concurrent branches (this stream + two auxiliary ones, forked / joined with events; inside a captured
  // graph these become parallel branches). b200tts_set_option("bigvgan_branches", 0) serialises them again.
  bool bigvgan_branches = true;
  // F5, ragged batches: the input embedding of ALL utterances as one launch over the 2 * Ntot DiT rows instead of two launches
  // per utterance (f5.cu): the profiled embedding time of configs[3] drops from 64 to 17 ms per batch, the step by 0.1-0.7 %
  // (the profile pass weighs short launches more than the graph replay does). b200tts_set_option("ragged_embed", 0) / B200TTS_RAGGED_EMBED=0
  // at engine creation restore the per-utterance launches.
  bool ragged_embed = true;
  cudaStream_t aux_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_acc[2] = {nullptr, nullptr}, ev_end[2] = {nullptr, nullptr};
  void ensure_aux();                       // creates the auxiliary streams / events (outside any capture)

  const Tensor& weight(const std::string& name) const {
    auto it = weights.find(name);
    if (it == weights.end()) fail("missing weight tensor: " + name);
    return it->second;
  }
  bool has_weight(const std::string& name) const { return weights.count(name) != 0; }
  ~Engine();
};

// arithmetic of the dense contractions; for the 16-bit engines the value doubles as the type code of 16-bit tensors
// (rowgemm.cuh out_bf16, aa_snake, ln_modulate): 1 = bf16, 2 = IEEE fp16
enum Precision : int { PREC_F32 = 0, PREC_BF16 = 1, PREC_F16 = 2 };

// Run `body` (a sequence of enqueues on e.stream, no host synchronisation inside) through the graph cache.
// `on_replay` restores whatever host-side state `body` would have set (it does not run when the graph is replayed).
template <typename Body, typename OnReplay>
void run_graphed(Engine& e, const std::vector<long long>& key, Body&& body, OnReplay&& on_replay) {
  if (!e.graphs.enabled || e.prof.enabled) { body(); return; }
  GraphCache::Entry* en = e.graphs.find(key);
  if (en == nullptr) {
    if (e.graphs.entries.size() >= 64) e.graphs.clear();
    body();
    GraphCache::Entry ne; ne.key = key; ne.epoch = g_alloc_epoch;
    e.graphs.entries.push_back(std::move(ne));
    return;
  }
  if (en->epoch != g_alloc_epoch) {           // some buffer moved since: the addresses baked into the graph may be stale
    if (en->exec) { cudaGraphExecDestroy(en->exec); en->exec = nullptr; }
    body();
    en = e.graphs.find(key);
    if (en) en->epoch = g_alloc_epoch;
    return;
  }
  if (en->exec == nullptr) {
    const unsigned long long l0 = g_launch_count;
    cudaGraph_t graph = nullptr;
    B2_CUDA(cudaStreamBeginCapture(e.stream, cudaStreamCaptureModeThreadLocal));
    try {
      body();
    } catch (...) {
      cudaStreamEndCapture(e.stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    B2_CUDA(cudaStreamEndCapture(e.stream, &graph));
    if (g_alloc_epoch != en->epoch) {           // body allocated while capturing: do not trust the capture, run eagerly
      cudaGraphDestroy(graph);
      en->epoch = g_alloc_epoch;
      body();
      return;
    }
    cudaGraphExec_t exec = nullptr;
    cudaError_t err = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) fail(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(err));
    en->exec = exec;
    en->launches = g_launch_count - l0;
    g_launch_count = l0;                        // counted per replay below
  }
  B2_CUDA(cudaGraphLaunch(en->exec, e.stream));
  g_launch_count += en->launches;
  on_replay();
}

}  // namespace b200tts

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

struct BigVGANModel;
struct F5Model;

struct Engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::map<std::string, Tensor> weights;   // reference state_dict names, prefixed "bigvgan." / "dit." / "vocos."
  Profiler prof;
  std::string prof_report;                 // last JSON report (owned here so the C ABI can hand out a pointer)
  BigVGANModel* bigvgan = nullptr;         // owned; freed by bigvgan_free / f5_free in ~Engine (api.cu)
  F5Model* f5 = nullptr;

  const Tensor& weight(const std::string& name) const {
    auto it = weights.find(name);
    if (it == weights.end()) fail("missing weight tensor: " + name);
    return it->second;
  }
  bool has_weight(const std::string& name) const { return weights.count(name) != 0; }
  ~Engine();
};

enum Precision : int { PREC_F32 = 0, PREC_BF16 = 1 };

}  // namespace b200tts

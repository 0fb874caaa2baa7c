// Shared helpers for libb200tts (sm_100a only).
#pragma once
#include <mutex>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace b200tts {

struct Error : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const std::string& msg) { throw Error(msg); }

#define B2_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::b200tts::fail(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                      __FILE__ + ":" + std::to_string(__LINE__));                          \
    }                                                                                       \
  } while (0)

#define B2_CHECK(cond, msg)                                                       \
  do {                                                                            \
    if (!(cond)) ::b200tts::fail(std::string("check failed: ") + #cond + ": " + (msg)); \
  } while (0)

#define B2_LAUNCH_CHECK() B2_CUDA(cudaGetLastError())

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
inline long round_up(long a, long b) { return (a + b - 1) / b * b; }

// Every kernel launch of ours goes through this counter (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

// Programmatic dependent launch. The hot path is thousands of short kernels in one stream (16 us each on average for one
// F5 utterance), so the launch / drain / fill gap between two of them is a visible share of the step. Every hot kernel
//   1. signals pdl_trigger() once its own on-chip resources are claimed (for tensor-memory users: AFTER tcgen05.alloc, so
//      that an early-arriving dependent CTA can never take the columns this CTA still waits for), which lets the next
//      kernel's CTAs be scheduled and run their prologue (barrier init, TMEM alloc, descriptor prefetch) as SMs drain, and
//   2. executes pdl_wait() before its first access to global memory: it returns when the preceding kernel has completed
//      and its writes are visible (a no-op for a kernel launched without the attribute).
// Both are no-ops semantically for ordering: the data dependence of a stream is unchanged.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();     // B200TTS_PDL=0 turns the launch attribute off (rowgemm_f32.cu)
// The attribute is only set inside a PdlScope(true): it pays where kernels are short (one-utterance DiT step: 77.2 -> 75.2 ms,
// the GPT decode chain) and costs ~0.8 % where they are long (BigVGAN 14.79 -> 15.04 ms, 8-utterance pipeline 444.7 -> 447.8 ms,
// same-box A/B), because an early-resident dependent CTA holds shared memory / TMEM that the tail of its predecessor could use.
extern int g_pdl_scope;
struct PdlScope {
  int prev;
  explicit PdlScope(bool on) : prev(g_pdl_scope) { g_pdl_scope = on ? 1 : 0; }
  ~PdlScope() { g_pdl_scope = prev; }
};

// Launch `kernel` so that it may start while its predecessor in the stream drains (its pdl_wait() orders the data).
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (pdl_enabled() && g_pdl_scope) ? 1 : 0;
  B2_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}
#endif

// Function attributes (cudaFuncSetAttribute) belong to the CURRENT DEVICE: a process that drives several GPUs (session.get_engine(d))
// must set them once per device, not once per process (ADVICE r01 flagged the same pattern in gpt2.cu). first() is true the first
// time it is called with a given device current.
struct PerDeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    std::lock_guard<std::mutex> g(mu);
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// Bumped by every device (re)allocation / free of a DevBuf: captured CUDA graphs bake buffer addresses into their nodes,
// so a cached graph is only replayed while the epoch it was captured under still holds (engine.cuh: GraphCache).
extern unsigned long long g_alloc_epoch;

// Device buffer with RAII.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) { B2_CUDA(cudaMalloc((void**)&p, count * sizeof(T))); ++g_alloc_epoch; }
  }
  // grow-only (never shrinks): workspaces are sized by the largest request seen
  void reserve(size_t count) { if (count > n) alloc(count); }
  void release() { if (p) { cudaFree(p); ++g_alloc_epoch; } p = nullptr; n = 0; }
};

// Epilogue activation selectors shared by the fp32 and tcgen05 GEMMs.
enum Act : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_MISH = 3 };

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case ACT_GELU_TANH: {   // torch GELU(approximate="tanh"): F5 modules.py:597
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      float u = k0 * (v + k1 * v * v * v);
      return 0.5f * v * (1.0f + tanhf(u));
    }
    case ACT_GELU_ERF:      // torch nn.GELU(): F5 modules.py:249, vocos modules.py:37
      return 0.5f * v * (1.0f + erff(v * 0.7071067811865476f));
    case ACT_MISH: {        // nn.Mish: x * tanh(softplus(x)), softplus threshold 20 (F5 modules.py:172)
      float sp = v > 20.0f ? v : log1pf(expf(v));
      return v * tanhf(sp);
    }
    default:
      return v;
  }
}

}  // namespace b200tts

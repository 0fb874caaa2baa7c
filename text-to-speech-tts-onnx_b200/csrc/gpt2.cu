// IndexTTS GPT-2 acoustic model on the device (reference: IndexTTS/Export_IndexTTS.py:203-289, graphs B-E; host loop
// IndexTTS/Inference_IndexTTS_ONNX.py:726-781).
//
// Two regimes share one KV cache:
//   * prefill (the first call of a sentence: conditioning latents + text rows + start token, ~100-250 rows): the row-GEMMs
//     of the DiT path (rowgemm_f32 / tcgen05 rowgemm_tc) with LayerNorm, attention and cache-scatter kernels around them;
//   * decode (one row per call, up to MAX_GENERATE_LENGTH calls): every projection is a matrix-VECTOR product whose cost is
//     the weight bytes it streams (24 layers x 12 D^2 + mel_head = 0.47 G parameters per token) -> HBM-bound. One warp per
//     output row, 16-byte loads, the LayerNorm that precedes a projection fused into its prologue (each CTA renormalises the
//     D-vector it is about to multiply), bias / gelu_new / residual / penalty in the epilogue. The loop state (cache length,
//     position, produced ids, penalty vector and its release window, stop flag) lives in device memory so that a decode step
//     is the SAME kernel sequence every time: it is captured once into a CUDA graph and replayed; the host looks at the stop
//     flag every few steps only.
// Layouts: hidden rows (rows, D) fp32; K and V caches [layer][head][row][64] fp32 (the reference keeps K transposed as
// (H, 64, S), gpt_kv_export produces that view); projection weights [N][K] (K contiguous) fp32 or bf16.
#include "gpt2.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "f5_kernels.cuh"
#include "layout.cuh"
#include "rowgemm.cuh"
#include "rowgemm_tc.cuh"

namespace b200tts {

namespace {

constexpr int HD = 64;
// device loop state
enum : int { ST_KV = 0, ST_GEN = 1, ST_N = 2, ST_STOP = 3, ST_RESET = 4, ST_LIMIT = 5, ST_ERR = 6, ST_WORDS = 8 };

struct LoopConst {
  int start_mel, stop_mel, range;
  float repeat_penalty;
};

}  // namespace

struct GptLayer {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *o_b, *fc_b, *p_b;
  const float *o_kn, *fc_kn, *p_kn;          // Hugging Face Conv1D weights are (in, out) = [K][N]: the fp32 GEMM layout as is
  DevBuf<float> qkv_b;                       // q and k thirds scaled by 64^-0.25 (Export_IndexTTS.py:250-255)
  DevBuf<float> qkv_kn;                      // [D][3D], scaled, for the fp32 prefill GEMM
  DevBuf<float> qkv_nk, o_nk, fc_nk, p_nk;   // [N][K] fp32: decode GEMV of the fp32 engine (built on first use)
  TcWeight qkv_tc, o_tc, fc_tc, p_tc;        // [N][K] bf16: prefill GEMM and decode GEMV of the bf16 engine
};

struct GptModel {
  int D = 0, L = 0, H = 0, FF = 0, Vm = 0, Vt = 0, Pt = 0, Pm = 0, S_max = 0;
  float eps = 1e-5f;
  LoopConst lc{};
  int start_text = 0, stop_text = 1;
  const float *text_emb, *text_pos, *mel_emb, *mel_pos, *lnf_w, *lnf_b, *fn_w, *fn_b, *head_b;
  const float* head_nk;                      // mel_head is an nn.Linear: (out, in) = [N][K] already
  TcWeight head_tc;
  std::vector<GptLayer> layers;
  bool f32_ready = false, bf16_ready = false;
  // cache + loop state
  DevBuf<float> kc, vc;                      // [L][H][S_max][64]
  DevBuf<int> state, ids;                    // ST_WORDS ; produced ids [S_max + 1]
  DevBuf<float> penalty, hid_save;           // [Vm] ; [S_max + 1][D]
  int resident = 0;                          // host mirror of ST_KV for the per-call (session) entry points
  // workspaces
  DevBuf<float> hp, nb32, qkv32, att32, ff32, hcur, logits;
  DevBuf<__nv_bfloat16> nb16, att16, ff16;
  DevBuf<int> idbuf;
  int* h_state = nullptr;                    // pinned
  DevBuf<unsigned char> players;             // device array of PLayer (persistent decode kernel)
  DevBuf<unsigned long long> trace;          // optional phase-timestamp trace (B200TTS_GPT_TRACE)
  DevBuf<unsigned long long> tagged;         // its {value, epoch} vectors: h | qkv | att | ff | logits
};

namespace {

GptModel& model(Engine& e) {
  if (!e.igpt) fail("IndexTTS GPT weights are not built (call b200tts_indextts_gpt_build)");
  return *e.igpt;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float gelu_new_f(float v) {      // Hugging Face NewGELUActivation (GPT2MLP.act)
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * v * (1.0f + tanhf(k0 * (v + k1 * v * v * v)));
}

template <typename WT> struct WVec;
template <> struct WVec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&w)[4]) {
    w[0] = __uint_as_float(v.x); w[1] = __uint_as_float(v.y); w[2] = __uint_as_float(v.z); w[3] = __uint_as_float(v.w);
  }
};
template <> struct WVec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&w)[8]) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[2 * i] = __uint_as_float(u[i] << 16);
      w[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
};

struct GemvArgs {
  const void* W; long ldw;                    // [N][ldw]
  const float* bias;                          // [N] or null
  const float* x; int K, N;
  const float *ln_w, *ln_b, *ln2_w, *ln2_b;   // PRE >= 1: LayerNorm(x); PRE == 2: a second LayerNorm on top (ln_f, final_norm)
  float eps;
  const float* res;                           // [N] or null: y += res
  const float* mul;                           // [N] or null: y *= mul   (repeat penalty)
  float* y;
  float* save; const int* state;              // PRE == 2: block 0 stores the FIRST LayerNorm's output at save[state[ST_N]][K]
  int act;                                    // 0 none, 1 gelu_new
};

// LayerNorm of src[K] into dst[K] (both shared): every warp derives the row statistics itself (two passes over shared memory,
// no block barrier) and normalises its eighth of the row. The caller synchronises before dst is read.
__device__ __forceinline__ void ln_rows_shared(const float* src, float* dst, int K, const float* lw, const float* lb, float eps,
                                               int warp, int lane) {
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += src[k];
  const float mean = warp_sum_f(s) / (float)K;
  float q = 0.f;
  for (int k = lane; k < K; k += 32) { const float d = src[k] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum_f(q) / (float)K + eps);
  const int per = (K + 7) / 8, k1 = min(K, (warp + 1) * per);
  for (int k = warp * per + lane; k < k1; k += 32) dst[k] = (src[k] - mean) * rstd * lw[k] + lb[k];
}

// y[n] = epilogue(W[n][:] . pre(x) + bias[n]); one warp per output row, 8 rows per CTA.
// The first NCH 16-byte chunks of the row per lane are fetched BEFORE the grid dependency is awaited: the weights do not depend
// on the preceding kernel, so under programmatic dependent launch this kernel's HBM stream runs while its predecessor is still
// computing, and the part that is serialised behind the predecessor is x (a few KB from L2), the LayerNorm and the FMAs.
template <typename WT, int PRE, int NCH>
__global__ void __launch_bounds__(256) gemv_kernel(const GemvArgs a) {
  extern __shared__ float xs[];               // K floats (PRE == 0) or 2 K floats (raw | normalised)
  constexpr int V = WVec<WT>::N;
  const int K = a.K, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x * 8 + warp;
  const bool live = n < a.N;
  const WT* wr = reinterpret_cast<const WT*>(a.W) + (long)(live ? n : 0) * a.ldw;
  pdl_trigger();
  uint4 wq[NCH > 0 ? NCH : 1];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int k0 = (c * 32 + lane) * V;
    wq[c] = k0 < K ? __ldg(reinterpret_cast<const uint4*>(wr + k0)) : make_uint4(0u, 0u, 0u, 0u);
  }
  pdl_wait();
  for (int k = tid; k < K; k += 256) xs[k] = a.x[k];
  __syncthreads();
  const float* xv = xs;
  if constexpr (PRE >= 1) {
    float* xn = xs + K;
    ln_rows_shared(xs, xn, K, a.ln_w, a.ln_b, a.eps, warp, lane);
    __syncthreads();
    xv = xn;
    if constexpr (PRE == 2) {
      if (blockIdx.x == 0 && a.save != nullptr) {
        float* dst = a.save + (long)a.state[ST_N] * K;
        for (int k = tid; k < K; k += 256) dst[k] = xn[k];
      }
      ln_rows_shared(xn, xs, K, a.ln2_w, a.ln2_b, a.eps, warp, lane);
      __syncthreads();
      xv = xs;
    }
  }
  if (!live) return;
  float acc0 = 0.f, acc1 = 0.f;
  auto fma_chunk = [&](const uint4& q, int k0) {
    float w[V];
    WVec<WT>::unpack(q, w);
    const float4* x4p = reinterpret_cast<const float4*>(xv + k0);
#pragma unroll
    for (int i = 0; i < V / 4; ++i) {
      const float4 x4 = x4p[i];
      acc0 = fmaf(w[4 * i], x4.x, acc0);
      acc1 = fmaf(w[4 * i + 1], x4.y, acc1);
      acc0 = fmaf(w[4 * i + 2], x4.z, acc0);
      acc1 = fmaf(w[4 * i + 3], x4.w, acc1);
    }
  };
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int k0 = (c * 32 + lane) * V;
    if (k0 < K) fma_chunk(wq[c], k0);
  }
#pragma unroll 4
  for (int k0 = (NCH * 32 + lane) * V; k0 < K; k0 += 32 * V) fma_chunk(__ldg(reinterpret_cast<const uint4*>(wr + k0)), k0);
  float v = warp_sum_f(acc0 + acc1);
  if (lane == 0) {
    if (a.bias) v += a.bias[n];
    if (a.act == 1) v = gelu_new_f(v);
    if (a.res) v += a.res[n];
    if (a.mul) v *= a.mul[n];
    a.y[n] = v;
  }
}

// Attention of `rows` new query rows against the cache. grid (H, rows), 256 threads.
// qkv [rows][3D] fp32 (q | k | v thirds, head-major inside a third). The cache already holds the keys / values of the new rows
// when rows > 1 (kv_scatter_kernel); for rows == 1 this CTA appends its own head's row first. Key range of row r: causal ->
// [0, hist + r], else [0, hist + rows). The reference adds -128 to masked scores instead of removing them
// (Export_IndexTTS.py:245,268); exp(-128 - max) underflows to exactly 0 in fp32, so skipping them is the same arithmetic.
// Scores: one key per thread (a 256-byte cache row each). P.V: 16 key groups x 16 float4 columns, partial sums through shared memory.
constexpr int ATT_NT = 256;
template <typename OutT>
__global__ void __launch_bounds__(ATT_NT) gpt_attn_kernel(const float* __restrict__ qkv, float* __restrict__ kc, float* __restrict__ vc,
                                                          const int* __restrict__ state, int S_max, int D, int H, int causal,
                                                          OutT* __restrict__ out) {
  extern __shared__ float sm[];               // scores [S_max] | q [64] | part [16][64] | red [16]
  pdl_trigger();
  pdl_wait();
  float* sc = sm;
  float* qs = sm + S_max;
  float* part = qs + HD;
  float* red = part + 16 * HD;
  const int h = blockIdx.x, r = blockIdx.y, rows = gridDim.y, tid = threadIdx.x;
  const int hist = state[ST_KV];
  const float* qrow = qkv + (long)r * 3 * D + h * HD;
  float* kh = kc + (long)h * S_max * HD;
  float* vh = vc + (long)h * S_max * HD;
  if (tid < HD) qs[tid] = qrow[tid];
  if (rows == 1 && tid < 2 * HD) {            // decode: append this head's new key / value row
    if (tid < HD) kh[(long)hist * HD + tid] = qrow[D + tid];
    else vh[(long)hist * HD + tid - HD] = qrow[2 * D + tid - HD];
  }
  __syncthreads();
  const int nk = causal ? hist + r + 1 : hist + rows;
  float mx = -3.0e38f;
  if (tid < nk) {
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = qs[d];
    for (int j = tid; j < nk; j += ATT_NT) {
      const float4* kr = reinterpret_cast<const float4*>(kh + (long)j * HD);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < HD / 4; ++i) {
        const float4 k4 = kr[i];
        s0 = fmaf(q[4 * i], k4.x, s0); s1 = fmaf(q[4 * i + 1], k4.y, s1);
        s0 = fmaf(q[4 * i + 2], k4.z, s0); s1 = fmaf(q[4 * i + 3], k4.w, s1);
      }
      const float s = s0 + s1;
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < ATT_NT / 32; ++i) mx = fmaxf(mx, red[i]);
  float sum = 0.f;
  for (int j = tid; j < nk; j += ATT_NT) { const float p = expf(sc[j] - mx); sc[j] = p; sum += p; }
  sum = warp_sum_f(sum);
  if ((tid & 31) == 0) red[8 + (tid >> 5)] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < ATT_NT / 32; ++i) tot += red[8 + i];
  const float inv = 1.0f / tot;
  // out[d] = sum_j p_j v[j][d]
  const int g = tid >> 4, dq = tid & 15;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int j = g; j < nk; j += 16) {
    const float p = sc[j];
    const float4 v4 = *reinterpret_cast<const float4*>(vh + (long)j * HD + dq * 4);
    o.x = fmaf(p, v4.x, o.x); o.y = fmaf(p, v4.y, o.y); o.z = fmaf(p, v4.z, o.z); o.w = fmaf(p, v4.w, o.w);
  }
  *reinterpret_cast<float4*>(part + g * HD + dq * 4) = o;
  __syncthreads();
  if (tid < HD) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) v += part[i * HD + tid];
    v *= inv;
    if constexpr (sizeof(OutT) == 4) out[(long)r * D + h * HD + tid] = v;
    else out[(long)r * D + h * HD + tid] = __float2bfloat16(v);
  }
}

// new rows of k / v (qkv thirds 2 and 3) -> cache rows [hist, hist + rows) of one layer
__global__ void kv_scatter_kernel(const float* __restrict__ qkv, float* __restrict__ kc, float* __restrict__ vc,
                                  const int* __restrict__ state, int rows, int S_max, int D) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)rows * D) return;
  const int r = (int)(i / D), c = (int)(i - (long)r * D);
  const int h = c / HD, d = c - h * HD;
  const int hist = state[ST_KV];
  const long dst = ((long)h * S_max + hist + r) * HD + d;
  kc[dst] = qkv[(long)r * 3 * D + D + c];
  vc[dst] = qkv[(long)r * 3 * D + 2 * D + c];
}

// argmax (first maximum, as torch.argmax) + the host loop's bookkeeping (Inference_IndexTTS_ONNX.py:757-781):
// record the id, advance the cache length by `rows`, stop on the stop id, else update the penalty window and write the next
// call's hidden row  mel_embedding[id] + mel_pos_embedding[gen_len]  (graph C).
__global__ void __launch_bounds__(1024) gpt_pick_kernel(const float* __restrict__ logits, int Vm, int* __restrict__ state,
                                                        int* __restrict__ ids, float* __restrict__ penalty,
                                                        const float* __restrict__ mel_emb, const float* __restrict__ mel_pos,
                                                        float* __restrict__ hcur, int D, int rows, LoopConst lc, int bookkeeping,
                                                        int* __restrict__ id_out) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  __shared__ int s_tok, s_go, s_gen;
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  float best = -3.0e38f; int idx = 0;      // idx stays a valid row even if no logit beats the sentinel (NaN / -inf rows)
  for (int n = tid; n < Vm; n += 1024) {
    const float v = logits[n];
    if (v > best) { best = v; idx = n; }       // ascending n per thread: keeps the first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  if ((tid & 31) == 0) { bv[tid >> 5] = best; bi[tid >> 5] = idx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w)
      if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
    s_tok = idx; s_go = 0; s_gen = 0;
    if (id_out) *id_out = idx;
    if (bookkeeping && !state[ST_STOP]) {
      const int n = state[ST_N];
      ids[n] = idx;
      state[ST_N] = n + 1;
      state[ST_KV] += rows;
      if (idx == lc.stop_mel) {
        state[ST_STOP] = 1;
      } else {
        penalty[idx] = lc.repeat_penalty;
        const int rs = state[ST_RESET];
        if (n + 1 > lc.range && ids[rs] != idx) { penalty[ids[rs]] = 1.0f; state[ST_RESET] = rs + 1; }
        s_gen = state[ST_GEN];
        state[ST_GEN] = s_gen + 1;
        s_go = 1;
        if (n + 1 >= state[ST_LIMIT]) state[ST_STOP] = 1;
      }
    }
  }
  __syncthreads();
  if (s_go) {
    const float* e = mel_emb + (long)s_tok * D;
    const float* p = mel_pos + (long)s_gen * D;
    for (int k = tid; k < D; k += 1024) hcur[k] = e[k] + p[k];
  }
}

// rows of  table[id] + pos[pos0 + r]  (graphs B and C; ids on the device). wrap != 0: ids are [start, ids..., stop]
__global__ void gpt_embed_kernel(const int* __restrict__ ids, int n_ids, const float* __restrict__ table, const float* __restrict__ pos,
                                 int pos0, float* __restrict__ out, int rows, int D, int wrap, int start_id, int stop_id, int vocab) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)rows * D) return;
  const int r = (int)(i / D), c = (int)(i - (long)r * D);
  int id;
  if (wrap) id = r == 0 ? start_id : (r == rows - 1 ? stop_id : ids[r - 1]);
  else id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  out[i] = table[(long)id * D + c] + pos[(long)(pos0 + r) * D + c];
}

__global__ void fill_kernel(float* x, long n, float v) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}
__global__ void scale_kernel(float* x, long n, float s) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}
// in [R][C] -> out [C][R]
__global__ void transpose32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < C) ? in[(long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(long)c * R + r] = t[threadIdx.x][i];
  }
}
// cache [H][S_max][64] -> key (H, 64, S), value (H, S, 64)
__global__ void kv_export_kernel(const float* __restrict__ kc, const float* __restrict__ vc, float* __restrict__ key,
                                 float* __restrict__ value, int H, int S, int S_max) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)H * S * HD) return;
  const int h = (int)(i / ((long)S * HD));
  const int rem = (int)(i - (long)h * S * HD);
  const int s = rem / HD, d = rem - s * HD;
  const long src = ((long)h * S_max + s) * HD + d;
  value[i] = vc[src];
  key[((long)h * HD + d) * S + s] = kc[src];
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent decode: ONE cooperative kernel runs whole decode calls (24 layers + head + pick, several tokens per launch).
// A decode call is a chain of 122 dependent matrix-vector phases of 3-13 MB each. As separate kernels (graph-replayed, with
// programmatic dependent launch) every link costs ~7 us against ~2 us of HBM time; a first persistent version with a grid
// barrier per phase was no faster (ncu: 51 % of the warp samples parked at the barrier, profiles/r01). So there is no barrier:
//   * every activation vector that crosses CTAs (h, qkv, att, ff, logits) is stored as 64-bit words {value, epoch}; a consumer
//     polls the words it needs until they carry the epoch of the phase it is in, so "wait for the producers" and "fetch the
//     vector" are the same L2 round trip, and no fence / atomic / flag sits between two phases;
//   * a CTA takes part in a phase only if it owns output rows there, hence every reader of a vector is also a producer of
//     the next one; since each output depends on the WHOLE input vector, a buffer can only be overwritten (one layer later)
//     after all of its readers are done -- write-after-read safety follows from the data flow itself;
//   * every warp issues the loads of the weight rows (and bias) it owns in its NEXT phase before it starts polling, so the
//     HBM stream keeps running while the dependent part (poll, LayerNorm in shared memory, FMAs, warp reduction) executes.
// The stop decision travels the same way: the pick phase tags the next token's h with a STOP epoch.
// Measured (B200TTS_GPT_TRACE + tools/igpt_trace.py, profiles/r01/igpt_phase_trace.md): 26.3 us per layer = ~11 us in the four
// store -> visible -> polled hand-offs (the poll's answer queues behind the weight rows the SM is ingesting), ~8 us issue-bound
// matrix-vector phases, 4.5 us attention (one CTA per head), ~2.7 us LayerNorms; the HBM time of a layer is 6 us. DESIGN.md 3b.
// ---------------------------------------------------------------------------------------------------------------------
typedef unsigned long long u64;
struct PLayer {
  const __nv_bfloat16 *wqkv, *wo, *wfc, *wp;           // [N][K] bf16
  const float *bqkv, *bo, *bfc, *bp, *ln1w, *ln1b, *ln2w, *ln2b;
};
struct PArgs {
  const PLayer* layers;
  int L, D, FF, H, Vm, S_max;
  float eps;
  const __nv_bfloat16* whead;
  const float *bhead, *lnfw, *lnfb, *fnw, *fnb, *mel_emb, *mel_pos;
  u64 *th, *tqkv, *tatt, *tff, *tlogits;               // tagged vectors
  float *hcur, *kc, *vc;
  int *state, *ids;
  float *penalty, *hid_save;
  int n_tokens, n0;                                    // tokens to attempt in this launch; tokens produced before it
  int poll_first;                                      // experiment (B200TTS_GPT_POLLFIRST=1): fetch a phase's weight rows AFTER its input
                                                       // was polled -- measured slower, 1158 vs 1375 tokens/s on one box; default 0
  unsigned long long* trace;                           // optional (B200TTS_GPT_TRACE): globaltimer stamps of token 1, [cta][layer][16]
  LoopConst lc;
};
constexpr int P_NT = 512, P_NW = P_NT / 32, P_SLOTS = 20, P_MAXR = 4;
constexpr unsigned EP_STOP = 0xffffffffu;

__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_tagged(u64* p, float v, unsigned ep) {
  const u64 w = ((u64)ep << 32) | (u64)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {       // weights: read once, keep them out of L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// slow path of a poll: the word did not carry `want` yet. abort: 1 = STOP epoch seen, 2 = timed out (~2 s)
__device__ __noinline__ u64 poll_slow(const u64* p, unsigned want, int* abort) {
  const long long t0 = clock64();
  u64 v = ld_relaxed_u64(p);
  while ((unsigned)(v >> 32) != want) {
    if ((unsigned)(v >> 32) == EP_STOP) { *abort = 1; break; }
    if (clock64() - t0 > 4000000000LL) { *abort = 2; break; }
    __nanosleep(40);
    v = ld_relaxed_u64(p);
  }
  return v;
}
// K tagged words -> floats in shared memory; all loads of a thread are issued before the first one is checked
template <int CE>
__device__ __forceinline__ void p_poll_x(float* xs, const u64* g, int K, unsigned want, int* abort) {
  u64 v[CE];
  __syncthreads();                                       // xs may still be read by the previous phase of slower warps
#pragma unroll
  for (int i = 0; i < CE; ++i) { const int k = threadIdx.x + i * P_NT; v[i] = k < K ? ld_relaxed_u64(g + k) : 0ull; }
#pragma unroll
  for (int i = 0; i < CE; ++i) {
    const int k = threadIdx.x + i * P_NT;
    if (k < K) {
      if ((unsigned)(v[i] >> 32) != want) v[i] = poll_slow(g + k, want, abort);
      xs[k] = __uint_as_float((unsigned)v[i]);
    }
  }
  __syncthreads();
}

template <int CH, int MAXR>
__device__ __forceinline__ void p_prefetch(uint4 (&wq)[P_SLOTS], float (&bq)[P_MAXR], const __nv_bfloat16* __restrict__ W,
                                           const float* __restrict__ bias, int N, int gw, int TW, int lane) {
  constexpr int K = CH * 256;
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    const int n = gw + i * TW;
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (i * CH + c < P_SLOTS)
        wq[i * CH + c] = n < N ? ldg_stream(W + (long)n * K + (c * 32 + lane) * 8) : make_uint4(0u, 0u, 0u, 0u);
    bq[i] = n < N ? __ldg(bias + n) : 0.f;
  }
}

template <int CH, int MAXR, typename Epi>
__device__ __forceinline__ void p_gemv(const uint4 (&wq)[P_SLOTS], const float (&bq)[P_MAXR], const __nv_bfloat16* __restrict__ W, int N,
                                       const float* xv, int gw, int TW, int lane, Epi&& epi) {
  constexpr int K = CH * 256;
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    const int n = gw + i * TW;
    if (n < N) {
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int k0 = (c * 32 + lane) * 8;
        uint4 q;
        if (i * CH + c < P_SLOTS) q = wq[i * CH + c];
        else q = ldg_stream(W + (long)n * K + k0);
        float w[8];
        WVec<__nv_bfloat16>::unpack(q, w);
        const float4 xa = *reinterpret_cast<const float4*>(xv + k0), xb = *reinterpret_cast<const float4*>(xv + k0 + 4);
        acc0 = fmaf(w[0], xa.x, acc0); acc1 = fmaf(w[1], xa.y, acc1); acc0 = fmaf(w[2], xa.z, acc0); acc1 = fmaf(w[3], xa.w, acc1);
        acc0 = fmaf(w[4], xb.x, acc0); acc1 = fmaf(w[5], xb.y, acc1); acc0 = fmaf(w[6], xb.z, acc0); acc1 = fmaf(w[7], xb.w, acc1);
      }
      const float v = warp_sum_f(acc0 + acc1);
      if (lane == 0) epi(n, v + bq[i]);
    }
  }
}

// LayerNorm src -> dst (shared): statistics per warp, one sixteenth of the row normalised per warp; ends with a block barrier
__device__ __forceinline__ void p_ln(const float* src, float* dst, int K, const float* lw, const float* lb, float eps, int warp, int lane) {
  // K % 128 == 0: each lane reads float4s, four independent partial sums (a first version with one scalar chain per lane spent
  // 1.7 us here: 2 x 40 dependent LDS + FADD)
  const float4* s4 = reinterpret_cast<const float4*>(src);
  const int n4 = K >> 2;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
  for (int k = lane; k < n4; k += 32) { const float4 v = s4[k]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
  const float mean = warp_sum_f((acc.x + acc.y) + (acc.z + acc.w)) / (float)K;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
  for (int k = lane; k < n4; k += 32) {
    const float4 v = s4[k];
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum_f((q.x + q.y) + (q.z + q.w)) / (float)K + eps);
  const int per = (K + P_NW - 1) / P_NW, k1 = min(K, (warp + 1) * per);
  for (int k = warp * per + lane; k < k1; k += 32) dst[k] = (src[k] - mean) * rstd * __ldg(lw + k) + __ldg(lb + k);
  __syncthreads();
}

// one head of the single new row against the cache (same arithmetic as gpt_attn_kernel with rows == 1), 512 threads
__device__ __forceinline__ void p_attention(const PArgs& a, int layer, int h, int hist, unsigned ep, float* sc, float* part, float* red,
                                            float* qs, int* abort) {
  const int tid = threadIdx.x, D = a.D, S_max = a.S_max;
  float* kh = a.kc + ((long)layer * a.H + h) * S_max * HD;
  float* vh = a.vc + ((long)layer * a.H + h) * S_max * HD;
  if (tid < 3 * HD) {                                    // q | k | v of this head, 64 words each
    const int part_id = tid >> 6, d = tid & (HD - 1);
    const u64* src = a.tqkv + part_id * D + h * HD + d;
    u64 v = ld_relaxed_u64(src);
    if ((unsigned)(v >> 32) != ep) v = poll_slow(src, ep, abort);
    const float f = __uint_as_float((unsigned)v);
    if (part_id == 0) qs[d] = f;
    else if (part_id == 1) kh[(long)hist * HD + d] = f;
    else vh[(long)hist * HD + d] = f;
  }
  __syncthreads();
  const int nk = hist + 1;
  // this thread's share of V (key group g, columns dq*4..): requested now, consumed after the softmax -- one DRAM round trip
  // for K and V together instead of two in sequence (the weight stream evicts the cache from L2 between tokens)
  constexpr int VPRE = 4;
  const int g = tid >> 4, dq = tid & 15;                   // 32 key groups x 16 float4 columns
  float4 vpre[VPRE];
#pragma unroll
  for (int i = 0; i < VPRE; ++i) {
    const int j = g + 32 * i;
    vpre[i] = j < nk ? __ldcg(reinterpret_cast<const float4*>(vh + (long)j * HD + dq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float mx = -3.0e38f;
  for (int j = tid; j < nk; j += P_NT) {
    const float4* kr = reinterpret_cast<const float4*>(kh + (long)j * HD);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      const float4 k4 = __ldcg(kr + i);
      const float4 q4 = *reinterpret_cast<const float4*>(qs + 4 * i);
      s0 = fmaf(q4.x, k4.x, s0); s1 = fmaf(q4.y, k4.y, s1); s0 = fmaf(q4.z, k4.z, s0); s1 = fmaf(q4.w, k4.w, s1);
    }
    const float s = s0 + s1;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < P_NW; ++i) mx = fmaxf(mx, red[i]);
  float sum = 0.f;
  for (int j = tid; j < nk; j += P_NT) { const float p = expf(sc[j] - mx); sc[j] = p; sum += p; }
  sum = warp_sum_f(sum);
  if ((tid & 31) == 0) red[P_NW + (tid >> 5)] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < P_NW; ++i) tot += red[P_NW + i];
  const float inv = 1.0f / tot;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < VPRE; ++i) {
    const int j = g + 32 * i;
    if (j < nk) {
      const float p = sc[j];
      o.x = fmaf(p, vpre[i].x, o.x); o.y = fmaf(p, vpre[i].y, o.y); o.z = fmaf(p, vpre[i].z, o.z); o.w = fmaf(p, vpre[i].w, o.w);
    }
  }
#pragma unroll 4
  for (int j = g + 32 * VPRE; j < nk; j += 32) {
    const float p = sc[j];
    const float4 v4 = __ldcg(reinterpret_cast<const float4*>(vh + (long)j * HD + dq * 4));
    o.x = fmaf(p, v4.x, o.x); o.y = fmaf(p, v4.y, o.y); o.z = fmaf(p, v4.z, o.z); o.w = fmaf(p, v4.w, o.w);
  }
  *reinterpret_cast<float4*>(part + g * HD + dq * 4) = o;
  __syncthreads();
  if (tid < HD) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) v += part[i * HD + tid];
    st_tagged(a.tatt + h * HD + tid, v * inv, ep);
  }
  __syncthreads();                                       // sc / part / red are reused by the next layer
}

// argmax over the tagged logits + the loop bookkeeping of gpt_pick_kernel, by CTA 0; publishes the next token's h
__device__ __forceinline__ void p_pick(const PArgs& a, unsigned ep_logits, unsigned ep_h_next, float* redv, int* redi, int* abort) {
  __shared__ int s_tok, s_go, s_gen;
  const int tid = threadIdx.x;
  float best = -3.0e38f; int idx = 0;      // idx stays a valid row even if no logit beats the sentinel (NaN / -inf rows)
  for (int n0 = 0; n0 < a.Vm; n0 += 8 * P_NT) {
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int n = n0 + i * P_NT + tid; v[i] = n < a.Vm ? ld_relaxed_u64(a.tlogits + n) : 0ull; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + i * P_NT + tid;
      if (n < a.Vm) {
        if ((unsigned)(v[i] >> 32) != ep_logits) v[i] = poll_slow(a.tlogits + n, ep_logits, abort);
        const float f = __uint_as_float((unsigned)v[i]);
        if (f > best) { best = f; idx = n; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  if ((tid & 31) == 0) { redv[tid >> 5] = best; redi[tid >> 5] = idx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < P_NW; ++w)
      if (redv[w] > best || (redv[w] == best && redi[w] < idx)) { best = redv[w]; idx = redi[w]; }
    int* state = a.state;
    s_tok = idx; s_go = 0; s_gen = 0;
    if (!*abort) {
      const int n = state[ST_N];
      a.ids[n] = idx;
      state[ST_N] = n + 1;
      state[ST_KV] += 1;
      if (idx == a.lc.stop_mel) {
        state[ST_STOP] = 1;
      } else {
        a.penalty[idx] = a.lc.repeat_penalty;
        const int rs = state[ST_RESET];
        if (n + 1 > a.lc.range && a.ids[rs] != idx) { a.penalty[a.ids[rs]] = 1.0f; state[ST_RESET] = rs + 1; }
        s_gen = state[ST_GEN];
        state[ST_GEN] = s_gen + 1;
        s_go = 1;
        if (n + 1 >= state[ST_LIMIT]) state[ST_STOP] = 1;
      }
    }
    __threadfence();                                     // ids / penalty / state before the tagged h that announces them
  }
  __syncthreads();
  __threadfence();
  const bool stop = a.state[ST_STOP] != 0 || *abort;
  const float* e = a.mel_emb + (long)s_tok * a.D;
  const float* p = a.mel_pos + (long)s_gen * a.D;
  for (int k = tid; k < a.D; k += P_NT) {
    const float hv = s_go ? __ldg(e + k) + __ldg(p + k) : 0.f;
    if (s_go) a.hcur[k] = hv;                            // plain copy for the host / the next launch
    st_tagged(a.th + k, hv, stop ? EP_STOP : ep_h_next);
  }
}

// CH_D = D / 256, CH_F = FF / 256; R_* = rows of that phase a warp may own (host checks ceil(N / warps) <= R_*)
template <int CH_D, int CH_F, int R_QKV, int R_FC, int R_HEAD>
__global__ void __launch_bounds__(P_NT, 1) gpt_decode_kernel(const PArgs a) {
  extern __shared__ float psm[];
  __shared__ int s_abort;
  const int D = a.D, FF = a.FF, L = a.L;
  constexpr int CE_D = (CH_D * 256 + P_NT - 1) / P_NT, CE_F = (CH_F * 256 + P_NT - 1) / P_NT;
  float* xs = psm;                       // FF
  float* xn = xs + FF;                   // D
  float* sc = xn + D;                    // S_max
  float* part = sc + a.S_max;            // 32 * 64
  float* red = part + 32 * HD;           // 2 * P_NW
  float* qs = red + 2 * P_NW;            // 64
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gw = blockIdx.x * P_NW + warp, TW = gridDim.x * P_NW;
  const int gw0 = blockIdx.x * P_NW;     // a CTA takes part in a phase iff it owns a row there (rows are dealt warp by warp)
  const bool in_qkv = gw0 < 3 * D, in_d = gw0 < D, in_fc = gw0 < FF, in_head = gw0 < a.Vm, in_att = (int)blockIdx.x < a.H;
  if (tid == 0) s_abort = 0;
  __syncthreads();
  uint4 wq[P_SLOTS];
  float bq[P_MAXR];
  float hreg = 0.f;
  const int hist0 = a.state[ST_KV];      // set by the host before the launch; +1 per token
  const bool pf = a.poll_first != 0;
  if (in_qkv && !pf) p_prefetch<CH_D, R_QKV>(wq, bq, a.layers[0].wqkv, a.layers[0].bqkv, 3 * D, gw, TW, lane);
#define STAMP(i) do { if (a.trace && t == 1 && tid == 0) a.trace[((size_t)blockIdx.x * L + l) * 16 + (i)] = gtimer(); } while (0)
  for (int t = 0; t < a.n_tokens; ++t) {
    const unsigned eh0 = 1u + (unsigned)t * (2u * L + 1u);
    for (int l = 0; l < L; ++l) {
      const PLayer& Ly = a.layers[l];
      const unsigned ev = 1u + (unsigned)(t * L + l);       // epoch of qkv / att / ff
      const unsigned eh = eh0 + 2u * l;                     // epoch of h entering the layer
      if (in_qkv) {
        // qkv = W_qkv . LN1(h) + b
        STAMP(0);
        p_poll_x<CE_D>(xs, a.th, D, eh, &s_abort);
        if (s_abort) goto done;
        if (pf) p_prefetch<CH_D, R_QKV>(wq, bq, Ly.wqkv, Ly.bqkv, 3 * D, gw, TW, lane);
        STAMP(1);
        if (lane == 0 && gw < D) hreg = xs[gw];            // residual operand of the rows this warp owns
        p_ln(xs, xn, D, Ly.ln1w, Ly.ln1b, a.eps, warp, lane);
        STAMP(2);
        p_gemv<CH_D, R_QKV>(wq, bq, Ly.wqkv, 3 * D, xn, gw, TW, lane, [&](int n, float v) { st_tagged(a.tqkv + n, v, ev); });
        STAMP(3);
      }
      if (in_att) { p_attention(a, l, blockIdx.x, hist0 + t, ev, sc, part, red, qs, &s_abort); STAMP(4); }
      if (in_d) {
        // h += W_o . att + b
        if (!pf) p_prefetch<CH_D, 1>(wq, bq, Ly.wo, Ly.bo, D, gw, TW, lane);
        STAMP(5);
        p_poll_x<CE_D>(xs, a.tatt, D, ev, &s_abort);
        if (s_abort) goto done;
        if (pf) p_prefetch<CH_D, 1>(wq, bq, Ly.wo, Ly.bo, D, gw, TW, lane);
        STAMP(6);
        p_gemv<CH_D, 1>(wq, bq, Ly.wo, D, xs, gw, TW, lane, [&](int n, float v) { st_tagged(a.th + n, v + hreg, eh + 1u); });
      }
      if (in_fc) {
        // ff = gelu_new(W_fc . LN2(h) + b)
        if (!pf) p_prefetch<CH_D, R_FC>(wq, bq, Ly.wfc, Ly.bfc, FF, gw, TW, lane);
        STAMP(7);
        p_poll_x<CE_D>(xs, a.th, D, eh + 1u, &s_abort);
        if (s_abort) goto done;
        if (pf) p_prefetch<CH_D, R_FC>(wq, bq, Ly.wfc, Ly.bfc, FF, gw, TW, lane);
        STAMP(8);
        if (lane == 0 && gw < D) hreg = xs[gw];
        p_ln(xs, xn, D, Ly.ln2w, Ly.ln2b, a.eps, warp, lane);
        STAMP(9);
        p_gemv<CH_D, R_FC>(wq, bq, Ly.wfc, FF, xn, gw, TW, lane, [&](int n, float v) { st_tagged(a.tff + n, gelu_new_f(v), ev); });
      }
      if (in_d) {
        // h += W_p . ff + b
        if (!pf) p_prefetch<CH_F, 1>(wq, bq, Ly.wp, Ly.bp, D, gw, TW, lane);
        STAMP(10);
        p_poll_x<CE_F>(xs, a.tff, FF, ev, &s_abort);
        if (s_abort) goto done;
        if (pf) p_prefetch<CH_F, 1>(wq, bq, Ly.wp, Ly.bp, D, gw, TW, lane);
        STAMP(11);
        p_gemv<CH_F, 1>(wq, bq, Ly.wp, D, xs, gw, TW, lane, [&](int n, float v) { st_tagged(a.th + n, v + hreg, eh + 2u); });
      }
      STAMP(12);
      if (l + 1 < L) {
        if (in_qkv && !pf) p_prefetch<CH_D, R_QKV>(wq, bq, a.layers[l + 1].wqkv, a.layers[l + 1].bqkv, 3 * D, gw, TW, lane);
      } else if (in_head) {
        if (!pf) p_prefetch<CH_D, R_HEAD>(wq, bq, a.whead, a.bhead, a.Vm, gw, TW, lane);
      }
    }
    if (in_head) {
      // head: ln_f (saved) -> final_norm -> mel_head * penalty
      p_poll_x<CE_D>(xs, a.th, D, eh0 + 2u * L, &s_abort);
      if (s_abort) goto done;
      if (pf) p_prefetch<CH_D, R_HEAD>(wq, bq, a.whead, a.bhead, a.Vm, gw, TW, lane);
      __threadfence();                                     // the penalty vector written by the previous pick
      p_ln(xs, xn, D, a.lnfw, a.lnfb, a.eps, warp, lane);
      if (blockIdx.x == 0) {
        float* dst = a.hid_save + (long)(a.n0 + t) * D;
        for (int k = tid; k < D; k += P_NT) dst[k] = xn[k];
      }
      p_ln(xn, xs, D, a.fnw, a.fnb, a.eps, warp, lane);
      p_gemv<CH_D, R_HEAD>(wq, bq, a.whead, a.Vm, xs, gw, TW, lane,
                           [&](int n, float v) { st_tagged(a.tlogits + n, v * __ldcg(a.penalty + n), 1u + (unsigned)t); });
    }
    if (in_qkv && !pf) p_prefetch<CH_D, R_QKV>(wq, bq, a.layers[0].wqkv, a.layers[0].bqkv, 3 * D, gw, TW, lane);
    if (blockIdx.x == 0) p_pick(a, 1u + (unsigned)t, eh0 + 2u * L + 1u, red, reinterpret_cast<int*>(red + P_NW), &s_abort);
  }
done:
  if (tid == 0 && s_abort == 2) atomicExch(a.state + ST_ERR, 1);
}

// tags of a new launch: h <- hcur with epoch 1, every other word epoch 0
__global__ void gpt_tag_init_kernel(u64* th, u64* tqkv, u64* tatt, u64* tff, u64* tlogits, const float* hcur, int D, int FF, int Vm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D) { th[i] = ((u64)1u << 32) | (u64)__float_as_uint(hcur[i]); tatt[i] = 0ull; }
  if (i < 3 * D) tqkv[i] = 0ull;
  if (i < FF) tff[i] = 0ull;
  if (i < Vm) tlogits[i] = 0ull;
}

#define LAUNCHED() do { B2_LAUNCH_CHECK(); count_launch(); } while (0)

void transpose_to(const float* in, float* out, int R, int C, cudaStream_t s) {
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
  transpose32_kernel<<<grid, block, 0, s>>>(in, out, R, C);
  B2_LAUNCH_CHECK();
}

const float* W(Engine& e, const std::string& name, long expect) {
  const Tensor& t = e.weight("igpt." + name);
  B2_CHECK(t.numel() == expect, "igpt." + name + ": unexpected size");
  return t.data.p;
}

// ---------------------------------------------------------------------------------------------------------------------
// weight layouts
// ---------------------------------------------------------------------------------------------------------------------
void prepare(Engine& e, GptModel& m, int precision) {
  const bool fast = precision == PREC_BF16;
  if (fast ? m.bf16_ready : m.f32_ready) return;
  cudaStream_t s = e.stream;
  const int D = m.D, FF = m.FF;
  DevBuf<float> tmp;
  if (fast) tmp.alloc((size_t)FF * D);
  for (int i = 0; i < m.L; ++i) {
    GptLayer& Ly = m.layers[i];
    struct Job { const float* kn; int K, N; DevBuf<float>* nk; TcWeight* tc; };
    const Job jobs[4] = {{Ly.qkv_kn.p, D, 3 * D, &Ly.qkv_nk, &Ly.qkv_tc}, {Ly.o_kn, D, D, &Ly.o_nk, &Ly.o_tc},
                         {Ly.fc_kn, D, FF, &Ly.fc_nk, &Ly.fc_tc}, {Ly.p_kn, FF, D, &Ly.p_nk, &Ly.p_tc}};
    for (const Job& j : jobs) {
      if (fast) {
        transpose_to(j.kn, tmp.p, j.K, j.N, s);                       // [K][N] -> [N][K]
        tc_weight_from_f32(*j.tc, tmp.p, 1, 1, j.N, j.K, s);
      } else {
        j.nk->alloc((size_t)j.N * j.K);
        transpose_to(j.kn, j.nk->p, j.K, j.N, s);
      }
    }
  }
  if (fast) {
    tc_weight_from_f32(m.head_tc, m.head_nk, 1, 1, m.Vm, D, s);
    std::vector<PLayer> pl(m.L);
    for (int i = 0; i < m.L; ++i) {
      GptLayer& Ly = m.layers[i];
      pl[i] = PLayer{Ly.qkv_tc.w.p, Ly.o_tc.w.p, Ly.fc_tc.w.p, Ly.p_tc.w.p, Ly.qkv_b.p, Ly.o_b, Ly.fc_b, Ly.p_b,
                     Ly.ln1_w, Ly.ln1_b, Ly.ln2_w, Ly.ln2_b};
    }
    m.players.alloc(pl.size() * sizeof(PLayer));
    B2_CUDA(cudaMemcpyAsync(m.players.p, pl.data(), pl.size() * sizeof(PLayer), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaStreamSynchronize(s));            // pl is a stack object
  }
  B2_CUDA(cudaStreamSynchronize(s));
  (fast ? m.bf16_ready : m.f32_ready) = true;
}

// ---------------------------------------------------------------------------------------------------------------------
// one call of graph E on the resident cache
// ---------------------------------------------------------------------------------------------------------------------
template <typename WT, int PRE>
void gemv(Engine& e, const char* tag, GemvArgs a) {
  const size_t smem = (size_t)a.K * sizeof(float) * (PRE >= 1 ? 2 : 1);
  B2_CHECK(a.K % 8 == 0 && a.ldw % 8 == 0 && smem <= 48 * 1024, "gemv: K must be a multiple of 8 and fit shared memory");
  const int chunks = ceil_div((long)a.K, 32 * WVec<WT>::N);      // 16-byte chunks of a weight row per lane
  ProfScope ps(e.prof, tag, e.stream);
  const dim3 grid(ceil_div(a.N, 8)), block(256);
  if (chunks >= 20) launch_pdl(gemv_kernel<WT, PRE, 20>, grid, block, smem, e.stream, a);
  else if (chunks >= 10) launch_pdl(gemv_kernel<WT, PRE, 10>, grid, block, smem, e.stream, a);
  else if (chunks >= 5) launch_pdl(gemv_kernel<WT, PRE, 5>, grid, block, smem, e.stream, a);
  else launch_pdl(gemv_kernel<WT, PRE, 2>, grid, block, smem, e.stream, a);
  LAUNCHED();
}

template <typename OutT>
void attention(Engine& e, GptModel& m, int layer, const float* qkv, int rows, int causal, OutT* out) {
  cudaStream_t s = e.stream;
  float* kc = m.kc.p + (size_t)layer * m.H * m.S_max * HD;
  float* vc = m.vc.p + (size_t)layer * m.H * m.S_max * HD;
  if (rows > 1) {
    ProfScope ps(e.prof, "igpt.kv_scatter", s);
    kv_scatter_kernel<<<ceil_div((long)rows * m.D, 256), 256, 0, s>>>(qkv, kc, vc, m.state.p, rows, m.S_max, m.D);
    LAUNCHED();
  }
  const size_t smem = (size_t)(m.S_max + HD + 16 * HD + 16) * sizeof(float);
  ProfScope ps(e.prof, "igpt.attention", s);
  launch_pdl(gpt_attn_kernel<OutT>, dim3(m.H, rows), dim3(ATT_NT), smem, s, qkv, kc, vc, (const int*)m.state.p, m.S_max, m.D, m.H, causal, out);
  LAUNCHED();
}

void reserve(GptModel& m, int rows, bool fast) {
  const size_t R = (size_t)rows;
  m.qkv32.reserve(R * 3 * m.D);
  m.hp.reserve(R * m.D);
  if (fast) { m.nb16.reserve(R * m.D); m.att16.reserve(R * m.D); m.ff16.reserve(R * m.FF); }
  m.nb32.reserve(R * m.D); m.att32.reserve(R * m.D); m.ff32.reserve(R * m.FF);
}

// Head: ln_f -> (saved) -> final_norm -> mel_head -> * penalty -> argmax (+ loop bookkeeping)
template <typename WT>
void head_and_pick(Engine& e, GptModel& m, const float* x_last, const float* penalty, int rows, int bookkeeping, float* save,
                   int* id_out) {
  GemvArgs a{};
  if constexpr (sizeof(WT) == 4) { a.W = m.head_nk; a.ldw = m.D; } else { a.W = m.head_tc.w.p; a.ldw = m.head_tc.ldc; }
  a.bias = m.head_b; a.x = x_last; a.K = m.D; a.N = m.Vm;
  a.ln_w = m.lnf_w; a.ln_b = m.lnf_b; a.ln2_w = m.fn_w; a.ln2_b = m.fn_b; a.eps = m.eps;
  a.mul = penalty; a.y = m.logits.p; a.save = save; a.state = m.state.p;
  gemv<WT, 2>(e, "igpt.head", a);
  ProfScope ps(e.prof, "igpt.pick", e.stream);
  launch_pdl(gpt_pick_kernel, dim3(1), dim3(1024), 0, e.stream, (const float*)m.logits.p, m.Vm, m.state.p, m.ids.p, m.penalty.p,
             m.mel_emb, m.mel_pos, m.hcur.p, m.D, rows, m.lc, bookkeeping, id_out);
  LAUNCHED();
}

// decode: one row (m.hcur) through the 24 layers with matrix-vector kernels
template <typename WT>
void decode_layers(Engine& e, GptModel& m) {
  const int D = m.D, FF = m.FF;
  for (int i = 0; i < m.L; ++i) {
    GptLayer& Ly = m.layers[i];
    auto wp = [&](DevBuf<float>& nk, TcWeight& tc, GemvArgs& a, int K) {
      if constexpr (sizeof(WT) == 4) { a.W = nk.p; a.ldw = K; } else { a.W = tc.w.p; a.ldw = tc.ldc; }
    };
    GemvArgs q{};
    wp(Ly.qkv_nk, Ly.qkv_tc, q, D);
    q.bias = Ly.qkv_b.p; q.x = m.hcur.p; q.K = D; q.N = 3 * D; q.ln_w = Ly.ln1_w; q.ln_b = Ly.ln1_b; q.eps = m.eps; q.y = m.qkv32.p;
    gemv<WT, 1>(e, "igpt.qkv_gemv", q);
    attention<float>(e, m, i, m.qkv32.p, 1, 0, m.att32.p);
    GemvArgs o{};
    wp(Ly.o_nk, Ly.o_tc, o, D);
    o.bias = Ly.o_b; o.x = m.att32.p; o.K = D; o.N = D; o.res = m.hcur.p; o.y = m.hcur.p;
    gemv<WT, 0>(e, "igpt.out_gemv", o);
    GemvArgs f{};
    wp(Ly.fc_nk, Ly.fc_tc, f, D);
    f.bias = Ly.fc_b; f.x = m.hcur.p; f.K = D; f.N = FF; f.ln_w = Ly.ln2_w; f.ln_b = Ly.ln2_b; f.eps = m.eps; f.act = 1; f.y = m.ff32.p;
    gemv<WT, 1>(e, "igpt.fc_gemv", f);
    GemvArgs p{};
    wp(Ly.p_nk, Ly.p_tc, p, FF);
    p.bias = Ly.p_b; p.x = m.ff32.p; p.K = FF; p.N = D; p.res = m.hcur.p; p.y = m.hcur.p;
    gemv<WT, 0>(e, "igpt.proj_gemv", p);
  }
}

// prefill: `rows` rows (m.hp) through the layers with row-GEMMs
void prefill_layers(Engine& e, GptModel& m, int rows, int causal, bool fast) {
  cudaStream_t s = e.stream;
  const int D = m.D, FF = m.FF;
  auto gemm = [&](const char* tag, const void* x, int K, int N, const float* kn, const TcWeight& tc, const float* bias, int act,
                  const float* res, void* out, int out_bf16) {
    RowGemm p;
    p.x = x; p.ldx = K; p.Lin = rows; p.Cin = K; p.N = N; p.taps = 1; p.M = rows; p.B = 1;
    p.out = out; p.ldo = N; p.out_bf16 = out_bf16; p.bias = bias; p.act = act; p.res = res;
    ProfScope ps(e.prof, tag, s);
    if (fast) rowgemm_tc(p, tc, s);
    else { p.w = kn; p.ldw = N; rowgemm_f32(p, s); }
  };
  for (int i = 0; i < m.L; ++i) {
    GptLayer& Ly = m.layers[i];
    {
      ProfScope ps(e.prof, "igpt.layernorm", s);
      if (fast) layernorm_affine_bf16(m.hp.p, Ly.ln1_w, Ly.ln1_b, m.nb16.p, rows, D, m.eps, s);
      else layernorm_affine(m.hp.p, Ly.ln1_w, Ly.ln1_b, m.nb32.p, rows, D, m.eps, s);
    }
    gemm("igpt.qkv_gemm", fast ? (const void*)m.nb16.p : (const void*)m.nb32.p, D, 3 * D, Ly.qkv_kn.p, Ly.qkv_tc, Ly.qkv_b.p, ACT_NONE,
         nullptr, m.qkv32.p, 0);
    if (fast) attention<__nv_bfloat16>(e, m, i, m.qkv32.p, rows, causal, m.att16.p);
    else attention<float>(e, m, i, m.qkv32.p, rows, causal, m.att32.p);
    gemm("igpt.out_gemm", fast ? (const void*)m.att16.p : (const void*)m.att32.p, D, D, Ly.o_kn, Ly.o_tc, Ly.o_b, ACT_NONE, m.hp.p,
         m.hp.p, 0);
    {
      ProfScope ps(e.prof, "igpt.layernorm", s);
      if (fast) layernorm_affine_bf16(m.hp.p, Ly.ln2_w, Ly.ln2_b, m.nb16.p, rows, D, m.eps, s);
      else layernorm_affine(m.hp.p, Ly.ln2_w, Ly.ln2_b, m.nb32.p, rows, D, m.eps, s);
    }
    gemm("igpt.fc_gemm", fast ? (const void*)m.nb16.p : (const void*)m.nb32.p, D, FF, Ly.fc_kn, Ly.fc_tc, Ly.fc_b, ACT_GELU_TANH, nullptr,
         fast ? (void*)m.ff16.p : (void*)m.ff32.p, fast ? 1 : 0);
    gemm("igpt.proj_gemm", fast ? (const void*)m.ff16.p : (const void*)m.ff32.p, FF, D, Ly.p_kn, Ly.p_tc, Ly.p_b, ACT_NONE, m.hp.p, m.hp.p,
         0);
  }
}

// Persistent decode (bf16 engine). Returns false when this model shape has no instantiation (the caller falls back to the
// per-kernel graph path).
template <int CH_D, int CH_F, int R_QKV, int R_FC, int R_HEAD>
void launch_persistent(Engine& e, GptModel& m, int n_tokens, int n0) {
  cudaStream_t s = e.stream;
  int dev = 0, sms = 0, coop = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B2_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  B2_CHECK(coop != 0, "persistent decode needs cooperative launch support");
  const int TW = sms * P_NW;
  B2_CHECK(ceil_div(3 * m.D, TW) <= R_QKV && ceil_div(m.FF, TW) <= R_FC && ceil_div(m.Vm, TW) <= R_HEAD && m.D <= TW,
           "persistent decode: too few SMs for the row ownership this instantiation assumes");
  auto kern = gpt_decode_kernel<CH_D, CH_F, R_QKV, R_FC, R_HEAD>;
  const size_t smem = (size_t)(m.FF + m.D + m.S_max + 32 * HD + 2 * P_NW + HD) * sizeof(float);
  // the opt-in shared-memory size is a per-device function attribute and depends on the model (S_max): raise it whenever this
  // (device, size) needs more than what was set before (a second model / a second GPU in one process)
  static std::mutex attr_mu;
  static std::map<int, size_t> attr_smem;
  {
    std::lock_guard<std::mutex> lk(attr_mu);
    size_t& cur = attr_smem[dev];
    if (smem > cur) {
      B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, P_NT, smem));
      B2_CHECK(per_sm >= 1, "persistent decode kernel does not fit an SM");
      cur = smem;
    }
  }
  PArgs pa{};
  pa.layers = reinterpret_cast<const PLayer*>(m.players.p);
  pa.L = m.L; pa.D = m.D; pa.FF = m.FF; pa.H = m.H; pa.Vm = m.Vm; pa.S_max = m.S_max; pa.eps = m.eps;
  pa.whead = m.head_tc.w.p; pa.bhead = m.head_b; pa.lnfw = m.lnf_w; pa.lnfb = m.lnf_b; pa.fnw = m.fn_w; pa.fnb = m.fn_b;
  pa.mel_emb = m.mel_emb; pa.mel_pos = m.mel_pos;
  const size_t nt = (size_t)m.D * 5 + m.FF + m.Vm;
  m.tagged.reserve(nt);
  pa.th = m.tagged.p; pa.tqkv = pa.th + m.D; pa.tatt = pa.tqkv + 3 * m.D; pa.tff = pa.tatt + m.D; pa.tlogits = pa.tff + m.FF;
  pa.hcur = m.hcur.p; pa.kc = m.kc.p; pa.vc = m.vc.p;
  pa.state = m.state.p; pa.ids = m.ids.p; pa.penalty = m.penalty.p; pa.hid_save = m.hid_save.p;
  pa.n_tokens = n_tokens; pa.n0 = n0; pa.lc = m.lc;
  { const char* v = getenv("B200TTS_GPT_POLLFIRST"); pa.poll_first = v != nullptr && atoi(v) != 0; }
  const char* trace_path = getenv("B200TTS_GPT_TRACE");
  const size_t trace_n = (size_t)sms * m.L * 16;
  if (trace_path && n0 == 1) {                              // the first decode launch of a sentence
    m.trace.reserve(trace_n);
    B2_CUDA(cudaMemsetAsync(m.trace.p, 0, trace_n * sizeof(unsigned long long), s));
    pa.trace = m.trace.p;
  }
  {
    const int big = std::max(std::max(3 * m.D, m.FF), m.Vm);
    gpt_tag_init_kernel<<<ceil_div(big, 256), 256, 0, s>>>(pa.th, pa.tqkv, pa.tatt, pa.tff, pa.tlogits, m.hcur.p, m.D, m.FF, m.Vm);
    LAUNCHED();
  }
  void* args[] = {&pa};
  ProfScope ps(e.prof, "igpt.decode_persistent", s);
  B2_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(sms), dim3(P_NT), args, smem, s));
  LAUNCHED();
  if (pa.trace) {
    std::vector<unsigned long long> h(trace_n);
    B2_CUDA(cudaMemcpyAsync(h.data(), m.trace.p, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), sizeof(unsigned long long), trace_n, f); fclose(f); }
  }
}

bool decode_persistent(Engine& e, GptModel& m, int n_tokens, int n0) {
  const char* v = getenv("B200TTS_GPT_PERSIST");
  if (v != nullptr && atoi(v) == 0) return false;
  if (m.head_tc.ldc != m.D) return false;
  if (m.D == 1280 && m.FF == 5120) { launch_persistent<5, 20, 2, 3, 4>(e, m, n_tokens, n0); return true; }
  if (m.D == 512 && m.FF == 2048 && m.Vm <= 2048) { launch_persistent<2, 8, 1, 1, 1>(e, m, n_tokens, n0); return true; }
  return false;
}

// One E call on the device state. x: rows > 1 -> m.hp holds the rows; rows == 1 -> m.hcur holds the row.
void e_call(Engine& e, GptModel& m, int rows, int causal, bool fast, const float* penalty, int bookkeeping, float* save, int* id_out) {
  PdlScope pdl(rows == 1);
  if (rows == 1) {
    if (fast) { decode_layers<__nv_bfloat16>(e, m); head_and_pick<__nv_bfloat16>(e, m, m.hcur.p, penalty, 1, bookkeeping, save, id_out); }
    else { decode_layers<float>(e, m); head_and_pick<float>(e, m, m.hcur.p, penalty, 1, bookkeeping, save, id_out); }
  } else {
    prefill_layers(e, m, rows, causal, fast);
    const float* last = m.hp.p + (size_t)(rows - 1) * m.D;
    if (fast) head_and_pick<__nv_bfloat16>(e, m, last, penalty, rows, bookkeeping, save, id_out);
    else head_and_pick<float>(e, m, last, penalty, rows, bookkeeping, save, id_out);
  }
}

void set_state(Engine& e, GptModel& m, int kv, int gen, int n, int stop, int reset, int limit) {
  int* h = m.h_state;
  for (int i = 0; i < ST_WORDS; ++i) h[i] = 0;
  h[ST_KV] = kv; h[ST_GEN] = gen; h[ST_N] = n; h[ST_STOP] = stop; h[ST_RESET] = reset; h[ST_LIMIT] = limit;
  B2_CUDA(cudaMemcpyAsync(m.state.p, h, ST_WORDS * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  B2_CUDA(cudaStreamSynchronize(e.stream));       // h_state is reused
}

}  // namespace

// =====================================================================================================================
// build
// =====================================================================================================================
GptModel* gpt_build(Engine& e) {
  std::unique_ptr<GptModel> mp(new GptModel());
  GptModel& m = *mp;
  cudaStream_t s = e.stream;
  const Tensor& te = e.weight("igpt.text_embedding.weight");
  const Tensor& me = e.weight("igpt.mel_embedding.weight");
  B2_CHECK(te.shape.size() == 2 && me.shape.size() == 2 && te.shape[1] == me.shape[1], "igpt embeddings: shape");
  m.D = (int)te.shape[1]; m.Vt = (int)te.shape[0]; m.Vm = (int)me.shape[0];
  B2_CHECK(m.D % 128 == 0 && m.D % HD == 0, "igpt: model_dim must be a multiple of 128");
  m.H = m.D / HD; m.FF = 4 * m.D;
  m.Pt = (int)e.weight("igpt.text_pos_embedding.emb.weight").shape[0];
  m.Pm = (int)e.weight("igpt.mel_pos_embedding.emb.weight").shape[0];
  const Tensor& mt = e.weight("igpt.meta");       // [start_text, stop_text, start_mel, stop_mel, max_generate, penalty_range, repeat_penalty, ln_eps]
  B2_CHECK(mt.numel() == 8, "igpt.meta must hold 8 values");
  float meta[8];
  B2_CUDA(cudaMemcpy(meta, mt.data.p, sizeof(meta), cudaMemcpyDeviceToHost));
  m.start_text = (int)meta[0]; m.stop_text = (int)meta[1];
  m.lc.start_mel = (int)meta[2]; m.lc.stop_mel = (int)meta[3];
  m.S_max = (int)meta[4]; m.lc.range = (int)meta[5]; m.lc.repeat_penalty = meta[6]; m.eps = meta[7];
  B2_CHECK(m.S_max >= 8 && m.S_max <= 8192, "igpt: max_generate out of range");
  B2_CHECK(m.lc.start_mel < m.Vm && m.lc.stop_mel < m.Vm, "igpt: start / stop mel ids outside the code book");
  while (e.has_weight("igpt.h." + std::to_string(m.L) + ".ln_1.weight")) ++m.L;
  B2_CHECK(m.L > 0, "igpt: no transformer layers loaded");
  const int D = m.D, FF = m.FF;
  m.text_emb = te.data.p; m.mel_emb = me.data.p;
  m.text_pos = W(e, "text_pos_embedding.emb.weight", (long)m.Pt * D);
  m.mel_pos = W(e, "mel_pos_embedding.emb.weight", (long)m.Pm * D);
  m.lnf_w = W(e, "ln_f.weight", D); m.lnf_b = W(e, "ln_f.bias", D);
  m.fn_w = W(e, "final_norm.weight", D); m.fn_b = W(e, "final_norm.bias", D);
  m.head_nk = W(e, "mel_head.weight", (long)m.Vm * D); m.head_b = W(e, "mel_head.bias", m.Vm);
  m.layers.resize(m.L);
  const float scale = powf((float)HD, -0.25f);
  for (int i = 0; i < m.L; ++i) {
    GptLayer& Ly = m.layers[i];
    const std::string p = "h." + std::to_string(i) + ".";
    Ly.ln1_w = W(e, p + "ln_1.weight", D); Ly.ln1_b = W(e, p + "ln_1.bias", D);
    Ly.ln2_w = W(e, p + "ln_2.weight", D); Ly.ln2_b = W(e, p + "ln_2.bias", D);
    Ly.o_kn = W(e, p + "attn.c_proj.weight", (long)D * D); Ly.o_b = W(e, p + "attn.c_proj.bias", D);
    Ly.fc_kn = W(e, p + "mlp.c_fc.weight", (long)D * FF); Ly.fc_b = W(e, p + "mlp.c_fc.bias", FF);
    Ly.p_kn = W(e, p + "mlp.c_proj.weight", (long)FF * D); Ly.p_b = W(e, p + "mlp.c_proj.bias", D);
    // q and k columns (and biases) x 64^-0.25 each (Export_IndexTTS.py:250-255): the scores come out divided by 8
    const float* qw = W(e, p + "attn.c_attn.weight", (long)D * 3 * D);
    const float* qb = W(e, p + "attn.c_attn.bias", 3 * D);
    Ly.qkv_b.alloc((size_t)3 * D);
    B2_CUDA(cudaMemcpyAsync(Ly.qkv_b.p, qb, (size_t)3 * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
    scale_kernel<<<ceil_div(2 * D, 256), 256, 0, s>>>(Ly.qkv_b.p, 2 * D, scale);
    // [D][3D] with the first 2D COLUMNS scaled: transpose -> scale the first 2D rows -> transpose back
    DevBuf<float> t((size_t)3 * D * D);
    transpose_to(qw, t.p, D, 3 * D, s);
    scale_kernel<<<ceil_div((long)2 * D * D, 256), 256, 0, s>>>(t.p, (long)2 * D * D, scale);
    Ly.qkv_kn.alloc((size_t)3 * D * D);
    transpose_to(t.p, Ly.qkv_kn.p, 3 * D, D, s);
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaStreamSynchronize(s));
  }
  const size_t kv = (size_t)m.L * m.H * m.S_max * HD;
  m.kc.alloc(kv); m.vc.alloc(kv);
  B2_CUDA(cudaMemsetAsync(m.kc.p, 0, kv * sizeof(float), s));
  B2_CUDA(cudaMemsetAsync(m.vc.p, 0, kv * sizeof(float), s));
  m.state.alloc(ST_WORDS); m.ids.alloc((size_t)m.S_max + 1);
  m.penalty.alloc((size_t)m.Vm); m.hid_save.alloc((size_t)(m.S_max + 1) * D);
  m.hcur.alloc((size_t)D); m.logits.alloc((size_t)m.Vm); m.idbuf.alloc(4);
  B2_CUDA(cudaMemsetAsync(m.ids.p, 0, ((size_t)m.S_max + 1) * sizeof(int), s));
  B2_CUDA(cudaMemsetAsync(m.state.p, 0, ST_WORDS * sizeof(int), s));
  B2_CUDA(cudaMallocHost((void**)&m.h_state, ST_WORDS * sizeof(int)));
  const size_t attn_smem = (size_t)(m.S_max + 17 * HD + 16) * sizeof(float);
  B2_CHECK(attn_smem <= 200 * 1024, "igpt: cache capacity too large for the attention kernel's score buffer");
  B2_CUDA(cudaFuncSetAttribute(gpt_attn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
  B2_CUDA(cudaFuncSetAttribute(gpt_attn_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
  B2_CUDA(cudaStreamSynchronize(s));
  return mp.release();
}

void gpt_free(GptModel* m) {
  if (!m) return;
  if (m->h_state) cudaFreeHost(m->h_state);
  delete m;
}
int gpt_dim(const GptModel& m) { return m.D; }
int gpt_layers(const GptModel& m) { return m.L; }
int gpt_heads(const GptModel& m) { return m.H; }
int gpt_mel_codes(const GptModel& m) { return m.Vm; }
int gpt_max_rows(const GptModel& m) { return m.S_max; }
int gpt_resident_rows(const GptModel& m) { return m.resident; }

// =====================================================================================================================
// graphs B, C
// =====================================================================================================================
void gpt_text_embed(Engine& e, const int* d_text_ids, int n_text, float* d_out) {
  GptModel& m = model(e);
  const int rows = n_text + 2;
  B2_CHECK(n_text >= 0 && rows <= m.Pt, "text_embed: more text ids than text_pos_embedding rows");
  gpt_embed_kernel<<<ceil_div((long)rows * m.D, 256), 256, 0, e.stream>>>(d_text_ids, n_text, m.text_emb, m.text_pos, 0, d_out, rows, m.D, 1,
                                                                          m.start_text, m.stop_text, m.Vt);
  LAUNCHED();
}

void gpt_mel_embed(Engine& e, const int* d_id, int gen_len, float* d_out) {
  GptModel& m = model(e);
  B2_CHECK(gen_len >= 0 && gen_len < m.Pm, "mel_embed: gen_len outside mel_pos_embedding");
  gpt_embed_kernel<<<ceil_div((long)m.D, 256), 256, 0, e.stream>>>(d_id, 1, m.mel_emb, m.mel_pos, gen_len, d_out, 1, m.D, 0, 0, 0, m.Vm);
  LAUNCHED();
}

// =====================================================================================================================
// graph E, per call (session surface)
// =====================================================================================================================
void gpt_step(Engine& e, const float* d_hidden, int rows, int history, int mask_flag, const float* d_penalty, int precision,
              float* d_last_hidden, int* d_max_id) {
  GptModel& m = model(e);
  B2_CHECK(precision == PREC_F32 || precision == PREC_BF16, "gpt_step: unknown precision");
  B2_CHECK(rows >= 1, "gpt_step: no rows");
  B2_CHECK(history == 0 || history == m.resident, "gpt_step: history_len does not match the resident KV cache");
  B2_CHECK(history + rows <= m.S_max, "gpt_step: KV cache capacity (MAX_GENERATE_LENGTH) exceeded");
  // The graph's mask is additive and sliced [:ids_len, :kv_seq_len] (Export_IndexTTS.py:245,272): with history > 0 AND several
  // new rows it would let row r see cache columns <= r only. The reference loop never does that (prefill has history 0, decode
  // has one row); this engine's causal path appends after the history, so the combination is refused rather than answered
  // differently.
  B2_CHECK(!(history > 0 && rows > 1 && mask_flag != 0),
           "gpt_step: a masked multi-row call on top of a resident history is not supported (the reference graph slices its "
           "mask from column 0; prefill with history_len = 0 or feed one row per call)");
  const bool fast = precision == PREC_BF16;
  prepare(e, m, precision);
  reserve(m, rows, fast);
  cudaStream_t s = e.stream;
  set_state(e, m, history, 0, 0, 0, 0, 1 << 30);
  float* dst = rows == 1 ? m.hcur.p : m.hp.p;
  B2_CUDA(cudaMemcpyAsync(dst, d_hidden, (size_t)rows * m.D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  // save slot 0 of hid_save receives ln_f(last row); bookkeeping off: the caller owns ids / penalty / positions
  e_call(e, m, rows, mask_flag != 0 ? 1 : 0, fast, d_penalty, 0, m.hid_save.p, d_max_id);
  B2_CUDA(cudaMemcpyAsync(d_last_hidden, m.hid_save.p, (size_t)m.D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  m.resident = history + rows;
}

void gpt_kv_export(Engine& e, int layer, float* d_key, float* d_value) {
  GptModel& m = model(e);
  B2_CHECK(layer >= 0 && layer < m.L, "kv_export: layer out of range");
  const int S = m.resident;
  if (S == 0) return;
  const size_t off = (size_t)layer * m.H * m.S_max * HD;
  kv_export_kernel<<<ceil_div((long)m.H * S * HD, 256), 256, 0, e.stream>>>(m.kc.p + off, m.vc.p + off, d_key, d_value, m.H, S, m.S_max);
  LAUNCHED();
}

// =====================================================================================================================
// one sentence, loop on the device
// =====================================================================================================================
int gpt_generate(Engine& e, const float* d_conds, int cond_rows, const int* d_text_ids, int n_text, int max_new, int precision,
                 float* d_penalty, int* d_ids_out, float* d_hidden_out) {
  GptModel& m = model(e);
  B2_CHECK(precision == PREC_F32 || precision == PREC_BF16, "gpt_generate: unknown precision");
  const bool fast = precision == PREC_BF16;
  cudaStream_t s = e.stream;
  const int D = m.D;
  const int rows = cond_rows + n_text + 2 + 1;             // graph D: conds | [start, text, stop] | first mel row
  B2_CHECK(cond_rows >= 0 && n_text >= 0 && rows >= 2, "gpt_generate: bad sizes");
  int limit = m.S_max - rows;                              // Inference_IndexTTS_ONNX.py:745
  if (max_new > 0 && max_new < limit) limit = max_new;
  B2_CHECK(limit >= 1, "gpt_generate: the prompt leaves no room to generate (MAX_GENERATE_LENGTH)");
  // positions run to `limit`: a checkpoint whose mel_pos_embedding is shorter than MAX_GENERATE_LENGTH (IndexTTS-1.0: 608 rows)
  // bounds the generation instead of failing calls that would never get that far
  if (limit + 1 > m.Pm) limit = m.Pm - 1;
  B2_CHECK(limit >= 1, "gpt_generate: mel_pos_embedding is empty");
  prepare(e, m, precision);
  reserve(m, rows, fast);
  if (d_penalty) {
    B2_CUDA(cudaMemcpyAsync(m.penalty.p, d_penalty, (size_t)m.Vm * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {                                                 // a fresh penalty vector (Inference_IndexTTS_ONNX.py:688)
    fill_kernel<<<ceil_div(m.Vm, 256), 256, 0, s>>>(m.penalty.p, m.Vm, 1.0f);
    LAUNCHED();
  }
  // graphs B, C (start mel id at position 0), D
  if (cond_rows) B2_CUDA(cudaMemcpyAsync(m.hp.p, d_conds, (size_t)cond_rows * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  gpt_text_embed(e, d_text_ids, n_text, m.hp.p + (size_t)cond_rows * D);
  {
    const int start = m.lc.start_mel;
    B2_CUDA(cudaMemcpyAsync(m.idbuf.p, &start, sizeof(int), cudaMemcpyHostToDevice, s));
    gpt_mel_embed(e, m.idbuf.p, 0, m.hp.p + (size_t)(rows - 1) * D);
  }
  set_state(e, m, /*kv*/ 0, /*gen*/ 1, /*n*/ 0, /*stop*/ 0, /*reset*/ 0, limit);
  m.resident = 0;
  // prefill: causal flag 1 (Inference:690), then single rows with flag 0 (:763-765)
  e_call(e, m, rows, 1, fast, m.penalty.p, 1, m.hid_save.p, nullptr);
  const int chunk = 32;
  int produced = 0, stopped = 0;
  const std::vector<long long> key = {30, precision, (long long)(uintptr_t)m.hcur.p, (long long)(uintptr_t)m.kc.p};
  while (!stopped) {
    B2_CUDA(cudaMemcpyAsync(m.h_state, m.state.p, ST_WORDS * sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    B2_CHECK(m.h_state[ST_ERR] == 0, "gpt_generate: the persistent decode kernel timed out at a grid barrier");
    produced = m.h_state[ST_N];
    stopped = m.h_state[ST_STOP];
    if (stopped) break;
    int todo = limit - produced;
    if (todo > chunk) todo = chunk;
    if (fast && decode_persistent(e, m, todo, produced)) continue;   // one cooperative kernel for the whole chunk
    for (int i = 0; i < todo; ++i)                          // steps after a stop inside the chunk leave the state untouched
      run_graphed(e, key, [&] { e_call(e, m, 1, 0, fast, m.penalty.p, 1, m.hid_save.p, nullptr); }, [] {});
  }
  m.resident = m.h_state[ST_KV];
  B2_CUDA(cudaMemcpyAsync(d_ids_out, m.ids.p, (size_t)produced * sizeof(int), cudaMemcpyDeviceToDevice, s));
  B2_CUDA(cudaMemcpyAsync(d_hidden_out, m.hid_save.p, (size_t)produced * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (d_penalty) B2_CUDA(cudaMemcpyAsync(d_penalty, m.penalty.p, (size_t)m.Vm * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return produced;
}

}  // namespace b200tts

// Fused anti-aliased SnakeBeta activation on channels-last tensors (B, L, C):
//   zero-pad | x2 polyphase up-FIR (12-tap Kaiser-sinc, gain 2) | x + sin^2(a x)/(b+1e-9) | zero-pad | 12-tap FIR, stride 2
// in ONE pass over HBM (the reference materialises two 2x-length intermediates).
// Reference: BigVGAN/modeling_modified/act.py:25-29, resample.py:30-34, filter.py:94-98, and the
// index -1 (15-sample) tables of bigvgan.py:370,381-382 for the POST variant (output L+30).
//
// Derivation (DESIGN.md "AA activation"): with x zero-extended and f the 12 taps,
//   u[2q]   = 2 * sum_{i<6} x[q-3+i] f[11-2i]        u[2q+1] = 2 * sum_{i<6} x[q-2+i] f[10-2i]
//   s[m]    = snake(u[m]) for m in [0, 2L) else 0                 (POST: no mask, natural tails)
//   out[t]  = sum_{j<12} f[j] s[2 t + j - 5],  t in [0, L)        (POST: t in [-15, L+15))
// Streaming form used here: the two new samples an output step appends, s[2t+7] and s[2t+8], are both FIRs over the SAME
// six inputs x[t+1..t+6]. A thread owns two adjacent channels (one 32-bit bf16x2 / 64-bit float2 access per row, so a
// warp touches 128 / 256 contiguous bytes) and walks a time segment with a 12-sample s window and a 6-sample x window
// per channel held in registers; the walk is unrolled by 6 so the window rotations are register renames, not moves.
// Per output: 12 + 12 FMA, 2 sin, ~6 snake ops -- against ~112 instructions per output of the first (tile-recompute)
// version (ncu, profiles/r01).
#include "aa_act.cuh"

#include <cuda_fp16.h>

namespace b200tts {

__constant__ float c_aa_f[12];
__constant__ float c_aa_f2[12];      // 2 * f (the up-sampler's gain folded into its taps)

void aa_set_filter(const float* taps12_host) {
  float f2[12];
  for (int i = 0; i < 12; ++i) f2[i] = 2.0f * taps12_host[i];
  B2_CUDA(cudaMemcpyToSymbol(c_aa_f, taps12_host, 12 * sizeof(float)));
  B2_CUDA(cudaMemcpyToSymbol(c_aa_f2, f2, 12 * sizeof(float)));
}

namespace {

constexpr int NT = 128;

// one row of a channel pair as loaded (kept packed while it waits in the prefetch queue: 1 register for bf16, 2 for fp32)
template <typename T> struct Raw;
template <> struct Raw<float> { float2 v; };
template <> struct Raw<__nv_bfloat16> { uint32_t v; };
template <> struct Raw<__half> { uint32_t v; };
__device__ __forceinline__ Raw<float> ldraw(const float* p, bool ok) {
  Raw<float> r; r.v = ok ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(0.f, 0.f); return r;
}
__device__ __forceinline__ Raw<__nv_bfloat16> ldraw(const __nv_bfloat16* p, bool ok) {
  Raw<__nv_bfloat16> r; r.v = ok ? __ldg(reinterpret_cast<const unsigned int*>(p)) : 0u; return r;
}
__device__ __forceinline__ Raw<__half> ldraw(const __half* p, bool ok) {
  Raw<__half> r; r.v = ok ? __ldg(reinterpret_cast<const unsigned int*>(p)) : 0u; return r;
}
__device__ __forceinline__ float2 unpack(const Raw<__half>& r) { return __half22float2(*reinterpret_cast<const __half2*>(&r.v)); }
__device__ __forceinline__ void st2(__half* p, float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack(const Raw<float>& r) { return r.v; }
__device__ __forceinline__ float2 unpack(const Raw<__nv_bfloat16>& r) {
  return make_float2(__uint_as_float(r.v << 16), __uint_as_float(r.v & 0xFFFF0000u));
}
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&v);
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2): the two channels a thread owns share every filter tap, so each FIR step
// is ONE instruction for both channels, with the tap as the broadcast scalar operand.
__device__ __forceinline__ unsigned long long pk2(float2 v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 up2(unsigned long long r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ float2 fma2(float2 a, float s, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(make_float2(s, s))), "l"(pk2(c)));
  return up2(d);
}
__device__ __forceinline__ float2 fma2v(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return up2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float s) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(make_float2(s, s))));
  return up2(d);
}
__device__ __forceinline__ float2 mul2v(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(d);
}

// One time segment [t0, t1) of one channel pair. EDGE = false is the interior fast path: every x row the segment touches
// (t0-5 .. t1+12) lies inside [0, L), every s sample inside [0, 2L) and t1 - t0 is a whole number of 6-step iterations,
// so the loads, the window updates and the stores carry no predicates (the bounds tests were a third of the issued
// instructions: ncu r01b, ALU pipe 38 % busy against FMA 28 %).
template <typename InT, typename OutT, bool PRECISE, bool POST, bool EDGE>
__device__ __forceinline__ void aa_segment(const InT* __restrict__ xb, OutT* __restrict__ yb, int C, int L, int t0, int t1,
                                           int t_first, float2 av, float2 ibv) {
  auto ldx = [&](int t) { return ldraw(xb + (long)t * C, !EDGE || (t >= 0 && t < L)); };
  // moving pointers for the main loop (one 64-bit bump per 6 steps; row k of the iteration is +k*C)
  const InT* xp = xb + (long)(t0 + 13) * C;        // first row prefetched by the first main iteration
  OutT* yp = yb + (long)(t0 - t_first) * C;
  auto snake = [&](float2 u) {                     // u + sin^2(a u) / (b + 1e-9), both channels
    const float2 arg = mul2v(u, av);
    float2 sn;
    sn.x = PRECISE ? sinf(arg.x) : __sinf(arg.x);
    sn.y = PRECISE ? sinf(arg.y) : __sinf(arg.y);
    return fma2v(mul2v(ibv, sn), sn, u);
  };

  // windows (logical index i lives in slot (i + rotation) % size, the rotation is a compile-time constant per step)
  float2 xw[6];      // x[t+1 .. t+6]   (.x / .y = the two channels)
  float2 w[12];      // s[2t-5 .. 2t+6]
#pragma unroll
  for (int i = 0; i < 12; ++i) w[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 6; ++i) xw[i] = unpack(ldx(t0 - 5 + i));

  // one streaming step at output time t (K = step number modulo 6): optionally emit out[t], then append s[2t+7], s[2t+8]
  // and x[t+7]. The 12-tap output FIR runs as two independent 6-tap chains (even / odd taps).
#define AA_STEP(K, t, emit)                                                                                          \
  {                                                                                                                  \
    if (emit) {                                                                                                      \
      float2 oa = mul2(w[(0 + 2 * K) % 12], c_aa_f[0]);                                                              \
      float2 ob = mul2(w[(1 + 2 * K) % 12], c_aa_f[1]);                                                              \
      _Pragma("unroll") for (int j = 2; j < 12; j += 2) {                                                            \
        oa = fma2(w[(j + 2 * K) % 12], c_aa_f[j], oa);                                                               \
        ob = fma2(w[(j + 1 + 2 * K) % 12], c_aa_f[j + 1], ob);                                                       \
      }                                                                                                              \
      const float2 o = add2(oa, ob);                                                                                 \
      st2(yp + (K) * (long)C, o.x, o.y);                                                                             \
    }                                                                                                                \
    float2 uo = mul2(xw[(0 + K) % 6], c_aa_f2[10]);                                                                  \
    float2 ue = mul2(xw[(0 + K) % 6], c_aa_f2[11]);                                                                  \
    _Pragma("unroll") for (int i = 1; i < 6; ++i) {                                                                  \
      uo = fma2(xw[(i + K) % 6], c_aa_f2[10 - 2 * i], uo);                                                           \
      ue = fma2(xw[(i + K) % 6], c_aa_f2[11 - 2 * i], ue);                                                           \
    }                                                                                                                \
    const bool ok_o = !EDGE || POST || ((t) >= -3 && (t) <= L - 4);     /* m = 2t+7 in [0, 2L) */                    \
    const bool ok_e = !EDGE || POST || ((t) >= -4 && (t) <= L - 5);     /* m = 2t+8 in [0, 2L) */                    \
    w[(0 + 2 * K) % 12] = ok_o ? snake(uo) : make_float2(0.f, 0.f);                                                  \
    w[(1 + 2 * K) % 12] = ok_e ? snake(ue) : make_float2(0.f, 0.f);                                                  \
    xw[(0 + K) % 6] = unpack(xn[K]);                            /* x[t+7], prefetched one iteration ahead */          \
  }

  // x rows are fetched six steps (one unrolled iteration) before they enter the window, so the DRAM latency is covered
  // by ~6 steps of math instead of one
  Raw<InT> xn[6], xf[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) xn[i] = ldx(t0 + 1 + i);         // consumed by the warm-up steps t0-6 .. t0-1
#pragma unroll
  for (int i = 0; i < 6; ++i) xf[i] = ldx(t0 + 7 + i);         // consumed by the first main iteration
  // warm-up: virtual steps t0-6 .. t0-1 fill the s window, no output
  AA_STEP(0, t0 - 6, false) AA_STEP(1, t0 - 5, false) AA_STEP(2, t0 - 4, false)
  AA_STEP(3, t0 - 3, false) AA_STEP(4, t0 - 2, false) AA_STEP(5, t0 - 1, false)
  for (int t = t0; t < t1; t += 6) {
#pragma unroll
    for (int i = 0; i < 6; ++i) xn[i] = xf[i];
    if (!EDGE || t + 6 < t1) {                   // (interior: the rows read past t1 exist, the values are simply unused)
#pragma unroll
      for (int i = 0; i < 6; ++i) xf[i] = ldraw(xp + i * (long)C, !EDGE || (t + 13 + i >= 0 && t + 13 + i < L));
    }
    xp += 6 * (long)C;
    AA_STEP(0, t + 0, true)
    AA_STEP(1, t + 1, (!EDGE || t + 1 < t1))
    AA_STEP(2, t + 2, (!EDGE || t + 2 < t1))
    AA_STEP(3, t + 3, (!EDGE || t + 3 < t1))
    AA_STEP(4, t + 4, (!EDGE || t + 4 < t1))
    AA_STEP(5, t + 5, (!EDGE || t + 5 < t1))
    yp += 6 * (long)C;
  }
#undef AA_STEP
}

template <typename InT, typename OutT, bool PRECISE, bool POST>
__global__ void __launch_bounds__(NT) aa_snake_kernel(const InT* __restrict__ x, OutT* __restrict__ y,
                                                      const float* __restrict__ alpha, const float* __restrict__ inv_beta,
                                                      int C, int L, long x_bstride, long y_bstride, int seg, int nseg) {
  pdl_trigger();
  pdl_wait();
  const int CP = C >> 1;
  const long gidx = (long)blockIdx.x * NT + threadIdx.x;
  const int sidx = (int)(gidx / CP);
  const int cp = (int)(gidx - (long)sidx * CP);
  if (sidx >= nseg) return;
  const int b = blockIdx.y;
  const int c = cp * 2;
  const int t_first = POST ? -15 : 0;
  const int t_end = POST ? L + 15 : L;
  const int t0 = t_first + sidx * seg;                     // first output time of this segment
  const int t1 = min(t0 + seg, t_end);
  const InT* xb = x + (long)b * x_bstride + c;
  OutT* yb = y + (long)b * y_bstride + c;
  const float2 av = make_float2(__ldg(alpha + c), __ldg(alpha + c + 1));
  const float2 ibv = make_float2(__ldg(inv_beta + c), __ldg(inv_beta + c + 1));
  // last iteration of an interior segment prefetches rows up to t1 + 12 (+6 more that are loaded but never used)
  const bool interior = t0 >= 5 && t1 + 18 < L && t1 - t0 == seg;
  if (interior) aa_segment<InT, OutT, PRECISE, POST, false>(xb, yb, C, L, t0, t1, t_first, av, ibv);
  else aa_segment<InT, OutT, PRECISE, POST, true>(xb, yb, C, L, t0, t1, t_first, av, ibv);
}

template <typename InT, typename OutT, bool PRECISE, bool POST>
void launch(const void* x, void* y, const float* alpha, const float* inv_beta, int B, int C, int L,
            cudaStream_t stream) {
  const int Lout = POST ? L + 30 : L;
  // segment length (multiple of 6): long enough to amortise the 6-step warm-up, short enough to fill the machine
  int seg = 96;
  while (seg > 24 && (long)B * (C / 2) * ceil_div(Lout, seg) < 250000L) seg -= 24;
  const int nseg = ceil_div(Lout, seg);
  const long threads = (long)nseg * (C / 2);
  dim3 grid(ceil_div(threads, NT), B);
  launch_pdl(aa_snake_kernel<InT, OutT, PRECISE, POST>, grid, dim3(NT), 0, stream,
             (const InT*)x, (OutT*)y, alpha, inv_beta, C, L, (long)L * C, (long)Lout * C, seg, nseg);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace

void aa_snake(const void* x, int in_bf16, void* y, int out_bf16, const float* alpha, const float* inv_beta,
              int B, int C, int L, bool precise, bool post, cudaStream_t stream) {
  B2_CHECK(B > 0 && C > 0 && L > 0, "aa_snake: empty tensor");
  B2_CHECK(C % 2 == 0, "aa_snake: channel count must be even");
  if (post) {
    B2_CHECK(!in_bf16 && !out_bf16, "aa_snake post variant is fp32 only");
    if (precise) launch<float, float, true, true>(x, y, alpha, inv_beta, B, C, L, stream);
    else launch<float, float, false, true>(x, y, alpha, inv_beta, B, C, L, stream);
    return;
  }
  if (precise) {
    B2_CHECK(!in_bf16 && !out_bf16, "aa_snake precise variant is fp32 only");
    launch<float, float, true, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (!in_bf16 && out_bf16 == 2) {                        // type codes: 0 = fp32, 1 = bf16, 2 = fp16
    launch<float, __half, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (in_bf16 == 2 && out_bf16 == 2) {
    launch<__half, __half, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (!in_bf16 && out_bf16 == 1) {
    launch<float, __nv_bfloat16, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (in_bf16 == 1 && out_bf16 == 1) {
    launch<__nv_bfloat16, __nv_bfloat16, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (!in_bf16 && !out_bf16) {
    launch<float, float, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else {
    fail("aa_snake: unsupported dtype combination");
  }
}

}  // namespace b200tts

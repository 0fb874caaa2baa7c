// Fused anti-aliased SnakeBeta activation on channels-last tensors (B, L, C):
//   zero-pad | x2 polyphase up-FIR (12-tap Kaiser-sinc, gain 2) | x + sin^2(a x)/(b+1e-9) | zero-pad | 12-tap FIR, stride 2
// in ONE pass over HBM (the reference materialises two 2x-length intermediates).
// Reference: BigVGAN/modeling_modified/act.py:25-29, resample.py:30-34, filter.py:94-98, and the
// index -1 (15-sample) tables of bigvgan.py:370,381-382 for the POST variant (output L+30).
//
// Derivation (DESIGN.md "AA activation"): with x zero-extended and f the 12 taps,
//   u[m]   = 2 * sum_i x[i] f[m + 5 - 2 i]                       (m even: taps 11,9,..,1; m odd: 10,8,..,0)
//   s[m]   = snake(u[m]) for m in [0, 2L) else 0                  (POST: no mask, natural tails)
//   out[t] = sum_{j<12} f[j] s[2 t + j - 5],  t in [0, L)         (POST: t in [-15, L+15))
// Each thread owns one channel and TT consecutive outputs, entirely in registers.
#include "aa_act.cuh"

namespace b200tts {

__constant__ float c_aa_f[12];

void aa_set_filter(const float* taps12_host) {
  B2_CUDA(cudaMemcpyToSymbol(c_aa_f, taps12_host, 12 * sizeof(float)));
}

namespace {

constexpr int TT = 16;
constexpr int NT = 256;

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename InT, typename OutT, bool PRECISE, bool POST>
__global__ void __launch_bounds__(NT) aa_snake_kernel(const InT* __restrict__ x, OutT* __restrict__ y,
                                                      const float* __restrict__ alpha,
                                                      const float* __restrict__ inv_beta, int C, int L,
                                                      long x_bstride, long y_bstride, int ntiles) {
  const long gidx = (long)blockIdx.x * NT + threadIdx.x;
  const int tile = (int)(gidx / C);
  const int c = (int)(gidx - (long)tile * C);
  if (tile >= ntiles) return;
  const int b = blockIdx.y;
  const int t_begin = POST ? -15 : 0;
  const int Lout = POST ? L + 30 : L;
  const int tb = t_begin + tile * TT;          // first logical output time of this tile

  const InT* xb = x + (long)b * x_bstride + c;
  float xw[TT + 10];                            // x[tb-5 .. tb+TT+4]
#pragma unroll
  for (int i = 0; i < TT + 10; ++i) {
    const int t = tb - 5 + i;
    xw[i] = (t >= 0 && t < L) ? ldf(xb + (long)t * C) : 0.f;
  }
  float f[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) f[j] = c_aa_f[j];
  const float a = __ldg(alpha + c), ib = __ldg(inv_beta + c);

  auto snake = [&](float u) {
    const float sn = PRECISE ? sinf(u * a) : __sinf(u * a);
    return fmaf(ib * sn, sn, u);
  };

  // s[sl] with sl = m - (2 tb - 5); ql = q - (tb - 3): odd m=2q+1 -> sl = 2 ql, even m=2q -> sl = 2 ql - 1
  float s[2 * TT + 10];
#pragma unroll
  for (int ql = 0; ql < TT + 6; ++ql) {
    if (ql >= 1) {               // even phase, x[q-3..q+2] -> xw[ql-1 .. ql+4], taps 11,9,7,5,3,1
      float u = xw[ql - 1] * f[11];
      u = fmaf(xw[ql + 0], f[9], u);
      u = fmaf(xw[ql + 1], f[7], u);
      u = fmaf(xw[ql + 2], f[5], u);
      u = fmaf(xw[ql + 3], f[3], u);
      u = fmaf(xw[ql + 4], f[1], u);
      u *= 2.0f;
      const int m = 2 * (tb - 3 + ql);
      const bool ok = POST || (m >= 0 && m < 2 * L);
      s[2 * ql - 1] = ok ? snake(u) : 0.f;
    }
    if (ql <= TT + 4) {          // odd phase, x[q-2..q+3] -> xw[ql .. ql+5], taps 10,8,6,4,2,0
      float u = xw[ql + 0] * f[10];
      u = fmaf(xw[ql + 1], f[8], u);
      u = fmaf(xw[ql + 2], f[6], u);
      u = fmaf(xw[ql + 3], f[4], u);
      u = fmaf(xw[ql + 4], f[2], u);
      u = fmaf(xw[ql + 5], f[0], u);
      u *= 2.0f;
      const int m = 2 * (tb - 3 + ql) + 1;
      const bool ok = POST || (m >= 0 && m < 2 * L);
      s[2 * ql] = ok ? snake(u) : 0.f;
    }
  }

  OutT* yb = y + (long)b * y_bstride + c;
#pragma unroll
  for (int r = 0; r < TT; ++r) {
    const int to = tile * TT + r;               // output row index (0-based in the output tensor)
    if (to < Lout) {
      float v = s[2 * r] * f[0];
#pragma unroll
      for (int j = 1; j < 12; ++j) v = fmaf(s[2 * r + j], f[j], v);
      stf(yb + (long)to * C, v);
    }
  }
}

template <typename InT, typename OutT, bool PRECISE, bool POST>
void launch(const void* x, void* y, const float* alpha, const float* inv_beta, int B, int C, int L,
            cudaStream_t stream) {
  const int Lout = POST ? L + 30 : L;
  const int ntiles = ceil_div(Lout, TT);
  const long threads = (long)ntiles * C;
  dim3 grid(ceil_div(threads, NT), B);
  aa_snake_kernel<InT, OutT, PRECISE, POST><<<grid, NT, 0, stream>>>(
      (const InT*)x, (OutT*)y, alpha, inv_beta, C, L, (long)L * C, (long)Lout * C, ntiles);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace

void aa_snake(const void* x, int in_bf16, void* y, int out_bf16, const float* alpha, const float* inv_beta,
              int B, int C, int L, bool precise, bool post, cudaStream_t stream) {
  B2_CHECK(B > 0 && C > 0 && L > 0, "aa_snake: empty tensor");
  if (post) {
    B2_CHECK(!in_bf16 && !out_bf16, "aa_snake post variant is fp32 only");
    if (precise) launch<float, float, true, true>(x, y, alpha, inv_beta, B, C, L, stream);
    else launch<float, float, false, true>(x, y, alpha, inv_beta, B, C, L, stream);
    return;
  }
  if (precise) {
    B2_CHECK(!in_bf16 && !out_bf16, "aa_snake precise variant is fp32 only");
    launch<float, float, true, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (!in_bf16 && out_bf16) {
    launch<float, __nv_bfloat16, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (in_bf16 && out_bf16) {
    launch<__nv_bfloat16, __nv_bfloat16, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else if (!in_bf16 && !out_bf16) {
    launch<float, float, false, false>(x, y, alpha, inv_beta, B, C, L, stream);
  } else {
    fail("aa_snake: unsupported dtype combination");
  }
}

}  // namespace b200tts

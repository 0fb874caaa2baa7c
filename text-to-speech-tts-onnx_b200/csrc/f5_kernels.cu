#include "f5_kernels.cuh"

#include <type_traits>

#include <cuda_fp16.h>
#include <cuda_fp8.h>

namespace b200tts {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per row, D <= 2048, D % 128 == 0. MODE 0: (1+a)*n + b ; MODE 1: a*n + b (affine LayerNorm)
// NV = D / 128 float4 per lane (compile-time, so the row and both modulation vectors sit in registers); the a / b loads are
// issued together with the row's, ahead of the two reductions, instead of after them (three dependent L2 round trips -> one).
template <int MODE, typename OutT, int NV>
__global__ void __launch_bounds__(256) rownorm_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                      const float* __restrict__ b, OutT* __restrict__ out, int R, int D,
                                                      float eps, float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long)row * D);
  float4 v[NV], aa[NV], bb[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + i * 32];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    aa[i] = __ldg(reinterpret_cast<const float4*>(a) + lane + i * 32);
    bb[i] = __ldg(reinterpret_cast<const float4*>(b) + lane + i * 32);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) sum += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    sq += dx * dx + dy * dy + dz * dz + dw * dw;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + eps);
  if (rstd_out != nullptr && lane == 0) rstd_out[row] = rstd;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * 32) * 4;
    float o[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
    const float av[4] = {aa[i].x, aa[i].y, aa[i].z, aa[i].w}, bv[4] = {bb[i].x, bb[i].y, bb[i].z, bb[i].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = MODE == 0 ? o[k] * (1.0f + av[k]) + bv[k] : o[k] * av[k] + bv[k];
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (long)row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
    } else if constexpr (std::is_same<OutT, __half>::value) {
      __half2 p0 = __floats2half2_rn(o[0], o[1]), p1 = __floats2half2_rn(o[2], o[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + (long)row * D + c) = pk;
    } else {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (long)row * D + c) = pk;
    }
  }
}

template <int MODE, typename OutT>
void launch_rownorm(const float* x, const float* a, const float* b, OutT* out, int R, int D, float eps, cudaStream_t s, float* rstd_out = nullptr) {
  const dim3 grid(ceil_div(R, 8));
  B2_CHECK(D % 128 == 0, "rownorm: row width must be a multiple of 128");
  switch (D / 128) {
    case 4: launch_pdl(rownorm_kernel<MODE, OutT, 4>, grid, dim3(256), 0, s, x, a, b, out, R, D, eps, rstd_out); break;     // text embedding (512)
    case 8: launch_pdl(rownorm_kernel<MODE, OutT, 8>, grid, dim3(256), 0, s, x, a, b, out, R, D, eps, rstd_out); break;     // DiT (1024)
    case 10: launch_pdl(rownorm_kernel<MODE, OutT, 10>, grid, dim3(256), 0, s, x, a, b, out, R, D, eps, rstd_out); break;   // IndexTTS GPT latent (1280)
    case 16: launch_pdl(rownorm_kernel<MODE, OutT, 16>, grid, dim3(256), 0, s, x, a, b, out, R, D, eps, rstd_out); break;
    default: fail("rownorm: row width must be 512, 1024, 1280 or 2048");
  }
}

__global__ void __launch_bounds__(256) l2norm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ b, float* __restrict__ out, int R, int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + (long)row * C;
  float sq = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = xr[c]; sq += v * v; }
  const float nrm = sqrtf(warp_sum(sq));
  for (int c = lane; c < C; c += 32) out[(long)row * C + c] = __ldg(w + c) * xr[c] / nrm + __ldg(b + c);
}

__global__ void dwconv7_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               float* __restrict__ out, int L, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)L * C) return;
  const int b = blockIdx.y;
  const int t = (int)(i / C), c = (int)(i - (long)t * C);
  const float* xb = x + (long)b * L * C;
  float acc = __ldg(bias + c);
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int tt = t + j - 3;
    if (tt >= 0 && tt < L) acc = fmaf(__ldg(w + j * C + c), xb[(long)tt * C + c], acc);
  }
  out[(long)b * L * C + i] = acc;
}

// block = (32 cols, 8 row lanes); gx[c] = sqrt(sum_r x[r][c]^2)
__global__ void __launch_bounds__(256) grn_colnorm_kernel(const float* __restrict__ x, float* __restrict__ gx, int R, int C) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (int r = threadIdx.y; r < R; r += 8) { const float v = x[(long)r * C + c]; s += v * v; }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    gx[c] = sqrtf(t);
  }
}
__global__ void __launch_bounds__(256) grn_mean_kernel(const float* __restrict__ gx, float* __restrict__ mean_out, int C) {
  __shared__ float red[8];
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) s += gx[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    *mean_out = t / (float)C;
  }
}
__global__ void grn_apply_kernel(float* __restrict__ x, const float* __restrict__ gx, const float* __restrict__ mean,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, long total, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const float nx = gx[c] / (*mean + 1e-6f);
  const float v = x[i];
  x[i] = __ldg(gamma + c) * (v * nx) + __ldg(beta + c) + v;
}

__global__ void text_gather_kernel(const int* __restrict__ ids, const float* __restrict__ table, const float* __restrict__ pos,
                                   float* __restrict__ out, int N, int D, int use_ids) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * D) return;
  const int n = (int)(i / D), d = (int)(i - (long)n * D);
  const int id = ids[n];
  out[i] = id == 0 ? 0.f : table[(long)(use_ids ? id : 0) * D + d] + pos[(long)n * D + d];
}
__global__ void mask_rows_kernel(float* __restrict__ x, const int* __restrict__ ids, int N, int D) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * D) return;
  if (ids[i / D] == 0) x[i] = 0.f;
}
__global__ void pad_ids_kernel(const int* __restrict__ text_ids, int n_text, int* __restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = i < n_text ? text_ids[i] + 1 : 0;
}
__global__ void reflect_pad_kernel(const int16_t* __restrict__ a, float* __restrict__ out, long L, int pad) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L + 2 * pad) return;
  long j = i - pad;
  if (j < 0) j = -j;
  if (j >= L) j = 2 * (L - 1) - j;
  out[i] = (float)a[j] * (1.0f / 32768.0f);
}
__global__ void magnitude_kernel(const float* __restrict__ spec, int ld, float* __restrict__ mag, int ldm, int F, int bins) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)F * ldm) return;
  const int f = (int)(i / ldm), c = (int)(i - (long)f * ldm);
  float v = 0.f;
  if (c < bins) {
    const float re = spec[(long)f * ld + c], im = spec[(long)f * ld + bins + c];
    v = sqrtf(re * re + im * im);
  }
  mag[i] = v;
}
__global__ void logmel_kernel(const float* __restrict__ mel, int F, float* __restrict__ dst, int ld, int col0, int N, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * C) return;
  const int n = (int)(i / C), c = (int)(i - (long)n * C);
  dst[(long)n * ld + col0 + c] = n < F ? logf(fmaxf(mel[i], 1e-5f)) : 0.f;
}
__global__ void copy_cols_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst, int col0, int N, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * C) return;
  const int n = (int)(i / C), c = (int)(i - (long)n * C);
  dst[(long)n * ld_dst + col0 + c] = src ? src[(long)n * ld_src + c] : 0.f;
}
__global__ void euler_kernel(float* __restrict__ noise, const float* __restrict__ pred, long n, float cfg, float dt) {
  pdl_trigger();
  pdl_wait();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long u = blockIdx.y;                       // utterance of the batch: pred is [U][2][n], noise [U][n]
  const float p0 = pred[(2 * u) * n + i], p1 = pred[(2 * u + 1) * n + i];
  noise[u * n + i] += (p0 + (p0 - p1) * cfg) * dt;
}
__global__ void istft_input_kernel(const float* __restrict__ head, float* __restrict__ out, int G, int bins, int ld) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)G * ld) return;
  const int g = (int)(i / ld), c = (int)(i - (long)g * ld);
  float v = 0.f;
  if (c < 2 * bins) {
    const int k = c < bins ? c : c - bins;
    const float mag = fminf(expf(head[(long)g * ld + k]), 100.0f);
    const float ph = head[(long)g * ld + bins + k];
    v = c < bins ? mag * cosf(ph) : mag * sinf(ph);
  }
  out[i] = v;
}
__global__ void overlap_add_kernel(const float* __restrict__ frames, const float* __restrict__ wsi, int G, int nfft, int hop,
                                   int16_t* __restrict__ pcm, float* __restrict__ wave) {
  const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n_out = (long)hop * (G - 1);
  if (o >= n_out) return;
  const long tp = o + nfft / 2;
  long f_lo = (tp - nfft + hop) / hop;        // ceil((tp - nfft + 1) / hop) for tp - nfft + 1 > 0
  if (tp - nfft + 1 <= 0) f_lo = 0;
  long f_hi = tp / hop;
  if (f_hi > G - 1) f_hi = G - 1;
  float acc = 0.f;
  for (long f = f_lo; f <= f_hi; ++f) acc += frames[f * nfft + (tp - f * hop)];
  float v = acc * wsi[tp];
  v = fminf(fmaxf(v, -1.0f), 1.0f) * 32767.0f;
  pcm[o] = (int16_t)v;
  if (wave) wave[o] = v;
}

// one thread per (row, head, pair): rope q and k (interleaved pairs), scatter kT and v
__global__ void rope_split_kernel(float* __restrict__ qkv, const float* __restrict__ cosT, const float* __restrict__ sinT,
                                  float* __restrict__ kT, float* __restrict__ v, int N, int H, int hd, int ldk) {
  const int D = H * hd;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = 2L * N * (D / 2);
  if (i >= total) return;
  const int pair = (int)(i % (D / 2));
  const long row = i / (D / 2);
  const int b = (int)(row / N), t = (int)(row % N);
  const int col = pair * 2, h = col / hd, d = col % hd;
  float* r = qkv + row * 3 * D;
  const float c0 = cosT[(long)t * hd + d], c1 = cosT[(long)t * hd + d + 1];
  const float s0 = sinT[(long)t * hd + d], s1 = sinT[(long)t * hd + d + 1];
  const float q0 = r[col], q1 = r[col + 1];
  r[col] = q0 * c0 + (-q1) * s0;             // x*cos + rotate_half(x)*sin with (x0,x1) -> (-x1, x0)
  r[col + 1] = q1 * c1 + q0 * s1;
  const float k0 = r[D + col], k1 = r[D + col + 1];
  const long kb = ((long)(b * H + h) * hd + d) * ldk + t;
  kT[kb] = k0 * c0 + (-k1) * s0;
  kT[kb + ldk] = k1 * c1 + k0 * s1;
  const long vb = ((long)(b * H + h) * ldk + t) * hd + d;      // v rows padded to ldk like kT's columns
  v[vb] = r[2 * D + col];
  v[vb + 1] = r[2 * D + col + 1];
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, long rows, int n, int ld) {
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* xr = x + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < n; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < n; c += 32) { const float e = expf(xr[c] - m); xr[c] = e; s += e; }
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int c = lane; c < ld; c += 32) xr[c] = c < n ? xr[c] * inv : 0.f;
}

__global__ void silu_kernel(const float* __restrict__ x, float* __restrict__ y, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; y[i] = v / (1.0f + expf(-v)); }
}

__global__ void rope_pack_kernel(const float* __restrict__ c, const float* __restrict__ sn, __half2* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __floats2half2_rn(c[i], sn[i]);
}

// u[t][j] = sum_k W[j][k] (1 + scale[t][k]),  v[t][j] = sum_k W[j][k] shift[t][k] + bias[j]   for the nfe rows of a modulation
// table, from the 16-bit weights the tensor cores multiply with (dit_chain.cu: LayerNorm folded into the GEMM). One block per j.
// Per-output-channel e4m3 quantisation of a weight [N][K] (fp32): scale[j] = max|W[j][:]| / 448, q[j][k] = e4m3(W[j][k] / scale[j])
__global__ void __launch_bounds__(128) quantize_rows_e4m3_kernel(const float* __restrict__ w, int K, uint8_t* __restrict__ q, int ldq,
                                                                 float* __restrict__ scale) {
  __shared__ float red[4];
  const int j = blockIdx.x, tid = threadIdx.x;
  float am = 0.f;
  for (int k = tid; k < K; k += 128) am = fmaxf(am, fabsf(w[(size_t)j * K + k]));
  am = warp_max(am);
  if ((tid & 31) == 0) red[tid >> 5] = am;
  __syncthreads();
  am = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  const float s = am > 0.f ? am / 448.0f : 1.0f;
  if (tid == 0) scale[j] = s;
  const float inv = 1.0f / s;
  for (int k = tid; k < K; k += 128)
    q[(size_t)j * ldq + k] = (uint8_t)__nv_cvt_float_to_fp8(w[(size_t)j * K + k] * inv, __NV_SATFINITE, __NV_E4M3);
}

// f16: 0 = bf16, 1 = fp16, 2 = e4m3 bytes with the per-row scale `wscale`
__global__ void __launch_bounds__(128) fold_vectors_kernel(const void* __restrict__ wv, const float* __restrict__ wscale, int ldc, int K, int f16,
                                                           const float* __restrict__ scale, const float* __restrict__ shift, int mod_ld,
                                                           const float* __restrict__ bias, float* __restrict__ u, float* __restrict__ v,
                                                           int N, int nfe) {
  __shared__ float red[2][4];
  const int j = blockIdx.x, tid = threadIdx.x;
  float wk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = tid + i * 128;
    float x = 0.f;
    if (k < K) {
      if (f16 == 2) {
        const __half_raw hr = __nv_cvt_fp8_to_halfraw(reinterpret_cast<const uint8_t*>(wv)[(size_t)j * ldc + k], __NV_E4M3);
        x = __half2float(__half(hr)) * wscale[j];
      } else {
        const uint16_t raw = reinterpret_cast<const uint16_t*>(wv)[(size_t)j * ldc + k];
        x = f16 ? __half2float(__ushort_as_half(raw)) : __bfloat162float(__ushort_as_bfloat16(raw));
      }
    }
    wk[i] = x;
  }
  for (int t = 0; t < nfe; ++t) {
    float su = 0.f, sv = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = tid + i * 128;
      if (k < K) {
        su = fmaf(wk[i], 1.0f + scale[(size_t)t * mod_ld + k], su);
        sv = fmaf(wk[i], shift[(size_t)t * mod_ld + k], sv);
      }
    }
    su = warp_sum(su); sv = warp_sum(sv);
    if ((tid & 31) == 0) { red[0][tid >> 5] = su; red[1][tid >> 5] = sv; }
    __syncthreads();
    if (tid == 0) {
      u[(size_t)t * N + j] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
      v[(size_t)t * N + j] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]) + bias[j];
    }
    __syncthreads();
  }
}

inline dim3 g1(long n, int bs = 256) { return dim3(ceil_div(n, bs)); }

}  // namespace

#define LAUNCHED() do { B2_LAUNCH_CHECK(); count_launch(); } while (0)

void ln_modulate(const float* x, const float* scale, const float* shift, void* out, int out_bf16, int R, int D, cudaStream_t s, float* rstd_out) {
  if (out_bf16 == 2) launch_rownorm<0, __half>(x, scale, shift, (__half*)out, R, D, 1e-6f, s, rstd_out);
  else if (out_bf16) launch_rownorm<0, __nv_bfloat16>(x, scale, shift, (__nv_bfloat16*)out, R, D, 1e-6f, s, rstd_out);
  else launch_rownorm<0, float>(x, scale, shift, (float*)out, R, D, 1e-6f, s, rstd_out);
  LAUNCHED();
}
void fold_vectors(const void* w16, int ldc, int K, int f16, const float* scale, const float* shift, int mod_ld, const float* bias, float* u,
                  float* v, int N, int nfe, cudaStream_t s, const float* wscale) {
  B2_CHECK(K <= 1024, "fold_vectors: K <= 1024");
  B2_CHECK(f16 != 2 || wscale != nullptr, "fold_vectors: e4m3 weights need their row scales");
  fold_vectors_kernel<<<N, 128, 0, s>>>(w16, wscale, ldc, K, f16, scale, shift, mod_ld, bias, u, v, N, nfe);
  LAUNCHED();
}
// e4m3 ff2 (dit_chain.cu, fp8 level 2): acc = sum_k e4m3(h_k * hgain) * e4m3(W_jk / sw_j), so gate * (acc * sw_j / hgain + b_j)
// = gate8 * (acc + bias8) with gate8[t][j] = gate[t][j] * sw_j / hgain and bias8[j] = b_j * hgain / sw_j
__global__ void fold_gate_bias_kernel(const float* __restrict__ gate, int mod_ld, const float* __restrict__ bias, const float* __restrict__ sw,
                                      float hgain, float* __restrict__ gate8, float* __restrict__ bias8, int N, int nfe) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const float s = sw[j] / hgain;
  bias8[j] = bias[j] / s;
  for (int t = 0; t < nfe; ++t) gate8[(size_t)t * N + j] = gate[(size_t)t * mod_ld + j] * s;
}
void fold_gate_bias(const float* gate, int mod_ld, const float* bias, const float* sw, float hgain, float* gate8, float* bias8, int N, int nfe,
                    cudaStream_t s) {
  fold_gate_bias_kernel<<<ceil_div(N, 128), 128, 0, s>>>(gate, mod_ld, bias, sw, hgain, gate8, bias8, N, nfe);
  LAUNCHED();
}
void quantize_rows_e4m3(const float* w, int N, int K, void* q, int ldq, float* scale, cudaStream_t s) {
  quantize_rows_e4m3_kernel<<<N, 128, 0, s>>>(w, K, reinterpret_cast<uint8_t*>(q), ldq, scale);
  LAUNCHED();
}
void layernorm_affine(const float* x, const float* w, const float* b, float* out, int R, int D, float eps, cudaStream_t s) {
  launch_rownorm<1, float>(x, w, b, out, R, D, eps, s);
  LAUNCHED();
}
void layernorm_affine_bf16(const float* x, const float* w, const float* b, __nv_bfloat16* out, int R, int D, float eps, cudaStream_t s) {
  launch_rownorm<1, __nv_bfloat16>(x, w, b, out, R, D, eps, s);
  LAUNCHED();
}
void l2_norm_affine(const float* x, const float* w, const float* b, float* out, int R, int C, cudaStream_t s) {
  l2norm_kernel<<<ceil_div(R, 8), 256, 0, s>>>(x, w, b, out, R, C);
  LAUNCHED();
}
void dwconv7(const float* x, const float* w, const float* bias, float* out, int B, int L, int C, cudaStream_t s) {
  dim3 grid(ceil_div((long)L * C, 256), B);
  dwconv7_kernel<<<grid, 256, 0, s>>>(x, w, bias, out, L, C);
  LAUNCHED();
}
void grn_inplace(float* x, const float* gamma, const float* beta, float* scratch, int R, int C, cudaStream_t s) {
  grn_colnorm_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, s>>>(x, scratch, R, C);
  LAUNCHED();
  grn_mean_kernel<<<1, 256, 0, s>>>(scratch, scratch + C, C);
  LAUNCHED();
  grn_apply_kernel<<<g1((long)R * C), 256, 0, s>>>(x, scratch, scratch + C, gamma, beta, (long)R * C, C);
  LAUNCHED();
}
void text_embed_gather(const int* ids, const float* table, const float* pos, float* out, int N, int D, int use_ids, cudaStream_t s) {
  text_gather_kernel<<<g1((long)N * D), 256, 0, s>>>(ids, table, pos, out, N, D, use_ids);
  LAUNCHED();
}
void mask_rows(float* x, const int* ids, int N, int D, cudaStream_t s) {
  mask_rows_kernel<<<g1((long)N * D), 256, 0, s>>>(x, ids, N, D);
  LAUNCHED();
}
void pad_text_ids(const int* text_ids, int n_text, int* ids_out, int N, cudaStream_t s) {
  pad_ids_kernel<<<g1(N), 256, 0, s>>>(text_ids, n_text, ids_out, N);
  LAUNCHED();
}
void audio_reflect_pad(const int16_t* audio, float* out, long L, int pad, cudaStream_t s) {
  B2_CHECK(L > pad, "audio shorter than the reflect padding");
  reflect_pad_kernel<<<g1(L + 2 * pad), 256, 0, s>>>(audio, out, L, pad);
  LAUNCHED();
}
void stft_magnitude(const float* spec, int ld, float* mag, int ldm, int F, int bins, cudaStream_t s) {
  magnitude_kernel<<<g1((long)F * ldm), 256, 0, s>>>(spec, ld, mag, ldm, F, bins);
  LAUNCHED();
}
void logmel_into(const float* mel, int F, float* dst, int ld_dst, int col0, int N, int C, cudaStream_t s) {
  logmel_kernel<<<g1((long)N * C), 256, 0, s>>>(mel, F, dst, ld_dst, col0, N, C);
  LAUNCHED();
}
void copy_cols(const float* src, int ld_src, float* dst, int ld_dst, int col0, int N, int C, cudaStream_t s) {
  copy_cols_kernel<<<g1((long)N * C), 256, 0, s>>>(src, ld_src, dst, ld_dst, col0, N, C);
  LAUNCHED();
}
void euler_cfg_update(float* noise, const float* pred, long n, int U, float cfg, float dt, cudaStream_t s) {
  launch_pdl(euler_kernel, dim3(ceil_div(n, 256), U), dim3(256), 0, s, noise, pred, n, cfg, dt);
  LAUNCHED();
}
void istft_input(const float* head, float* out, int G, int bins, int ld, cudaStream_t s) {
  istft_input_kernel<<<g1((long)G * ld), 256, 0, s>>>(head, out, G, bins, ld);
  LAUNCHED();
}
void istft_overlap_add(const float* frames, const float* wsi, int G, int nfft, int hop, int16_t* pcm, float* wave, cudaStream_t s) {
  if (G <= 1) return;
  overlap_add_kernel<<<g1((long)hop * (G - 1)), 256, 0, s>>>(frames, wsi, G, nfft, hop, pcm, wave);
  LAUNCHED();
}
void rope_split_f32(float* qkv, const float* cos, const float* sin, float* kT, float* v, int N, int H, int hd, int ldk, cudaStream_t s) {
  rope_split_kernel<<<g1(2L * N * (H * hd / 2)), 256, 0, s>>>(qkv, cos, sin, kT, v, N, H, hd, ldk);
  LAUNCHED();
}
void softmax_rows(float* x, long rows, int n, int ld, cudaStream_t s) {
  softmax_rows_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, rows, n, ld);
  LAUNCHED();
}
void rope_pack_half(const float* cos, const float* sin, __half2* out, long n, cudaStream_t s) {
  rope_pack_kernel<<<g1(n), 256, 0, s>>>(cos, sin, out, n);
  LAUNCHED();
}
void silu(const float* x, float* y, long n, cudaStream_t s) {
  silu_kernel<<<g1(n), 256, 0, s>>>(x, y, n);
  LAUNCHED();
}

}  // namespace b200tts

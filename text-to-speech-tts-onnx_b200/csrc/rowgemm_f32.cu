// SIMT fp32 shifted-row GEMM (see rowgemm.cuh): the exact-parity engine. 128x64 output tile per
// 256-thread CTA, 8x4 register tile per thread, BK=16, double-buffered shared memory.
#include "rowgemm.cuh"

namespace b200tts {

unsigned long long g_launch_count = 0;
unsigned long long g_alloc_epoch = 0;
int g_pdl_scope = 0;
bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("B200TTS_PDL"); return v == nullptr || atoi(v) != 0; }();
  return on;
}

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int LDA = BM + 4;   // As[k][m], padded, keeps float4 alignment

struct KArgs {
  const float* x; long x_bstride; int ldx; int Lin;
  const float* w; int ldw;
  int Cin, N, taps, dil, center, groups, M;
  void* out; long o_bstride; int ldo; long o_shift; long o_limit; int out_bf16;
  const float* bias; const float* gate; const float* res; int accumulate; float scale; int act;
  int chunks_per_tap;
};

__global__ void __launch_bounds__(NT) rowgemm_f32_kernel(const KArgs a) {
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int b = blockIdx.z / a.groups, g = blockIdx.z % a.groups;
  const int t0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  const float* xg = a.x + (long)b * a.x_bstride + (long)g * a.Cin;
  const float* wg = a.w + (long)g * a.taps * a.Cin * a.ldw;

  // A loader: thread -> (row = tid % 128, k-half = tid / 128): two float4 (8 consecutive channels)
  const int a_row = tid & (BM - 1), a_kh = tid >> 7;
  // B loader: thread -> (k = tid / 16, n4 = tid % 16): one float4
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  const int nchunks = a.taps * a.chunks_per_tap;
  float4 ra[2], rb;

  auto gload = [&](int q) {
    const int j = q / a.chunks_per_tap;
    const int c0 = (q - j * a.chunks_per_tap) * BK;
    const int t = t0 + a_row + (j - a.center) * a.dil;
    const bool row_ok = (t >= 0) && (t < a.Lin);
    const float* src = xg + (long)t * a.ldx + c0 + a_kh * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + a_kh * 8 + h * 4;
      ra[h] = (row_ok && c < a.Cin) ? __ldg(reinterpret_cast<const float4*>(src + h * 4))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int ck = c0 + b_k, n = n0 + b_n;
    rb = (ck < a.Cin && n < a.N)
             ? __ldg(reinterpret_cast<const float4*>(wg + ((long)j * a.Cin + ck) * a.ldw + n))
             : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = a_kh * 8 + h * 4;
      As[buf][k + 0][a_row] = ra[h].x;
      As[buf][k + 1][a_row] = ra[h].y;
      As[buf][k + 2][a_row] = ra[h].z;
      As[buf][k + 3][a_row] = ra[h].w;
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
  };

  const int tx = tid & 15, ty = tid >> 4;   // cols tx*4..+3, rows ty*8..+7
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();

  for (int q = 0; q < nchunks; ++q) {
    const int buf = q & 1;
    if (q + 1 < nchunks) gload(q + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    if (q + 1 < nchunks) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const int n = n0 + tx * 4;
  if (n >= a.N) return;
  const int gn = g * a.N + n;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), gate4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (a.bias) bias4 = __ldg(reinterpret_cast<const float4*>(a.bias + gn));
  if (a.gate) gate4 = __ldg(reinterpret_cast<const float4*>(a.gate + gn));
  const long obase = (long)b * a.o_bstride;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = t0 + ty * 8 + i;
    if (t >= a.M) break;
    const long flat = (long)t * a.ldo + gn + a.o_shift;
    if (flat < 0 || flat >= a.o_limit) continue;
    float v[4] = {acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w};
    if (a.act != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = act_apply(v[j], a.act);
    }
    v[0] *= gate4.x; v[1] *= gate4.y; v[2] *= gate4.z; v[3] *= gate4.w;
    if (a.res) {
      const float4 r = *reinterpret_cast<const float4*>(a.res + obase + flat);
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    if (a.out_bf16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + obase + flat;
      if (a.accumulate) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += __bfloat162float(o[j]);
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0] * a.scale, v[1] * a.scale);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2] * a.scale, v[3] * a.scale);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(o) = pk;
    } else {
      float* o = reinterpret_cast<float*>(a.out) + obase + flat;
      if (a.accumulate) {
        const float4 r = *reinterpret_cast<const float4*>(o);
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
      *reinterpret_cast<float4*>(o) = make_float4(v[0] * a.scale, v[1] * a.scale, v[2] * a.scale, v[3] * a.scale);
    }
  }
}

}  // namespace

void rowgemm_f32(const RowGemm& p, cudaStream_t stream) {
  B2_CHECK(p.Cin % 4 == 0 && p.N % 4 == 0, "rowgemm_f32 needs Cin, N multiples of 4");
  B2_CHECK(p.ldx % 4 == 0 && p.ldw % 4 == 0 && p.ldo % 4 == 0 && p.o_shift % 4 == 0, "rowgemm_f32 alignment");
  B2_CHECK(p.M > 0 && p.B > 0 && p.groups > 0 && p.taps > 0, "rowgemm_f32 empty problem");
  KArgs a;
  a.x = (const float*)p.x; a.x_bstride = p.x_bstride; a.ldx = p.ldx; a.Lin = p.Lin;
  a.w = (const float*)p.w; a.ldw = p.ldw;
  a.Cin = p.Cin; a.N = p.N; a.taps = p.taps; a.dil = p.dil; a.center = p.center; a.groups = p.groups; a.M = p.M;
  a.out = p.out; a.o_bstride = p.o_bstride; a.ldo = p.ldo; a.o_shift = p.o_shift;
  a.o_limit = p.o_limit ? p.o_limit : (long)p.M * p.ldo;
  a.out_bf16 = p.out_bf16;
  a.bias = p.bias; a.gate = p.gate; a.res = p.res; a.accumulate = p.accumulate; a.scale = p.scale; a.act = p.act;
  a.chunks_per_tap = ceil_div(p.Cin, BK);
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), p.B * p.groups);
  B2_CHECK(grid.y <= 65535 && grid.z <= 65535, "rowgemm_f32 grid too large");
  rowgemm_f32_kernel<<<grid, NT, 0, stream>>>(a);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts

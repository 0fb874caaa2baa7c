// Shifted-row GEMM on the 5th-gen tensor cores (sm_100a): TMA -> 128B-swizzled shared memory ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
//
//   D[128 rows (time) x BN (out channels)] += A[128 x 64] . B[BN x 64]^T     per (tap, 64-channel chunk)
//
// A tiles are boxes {64 ch, 128 rows, 1 batch} of the channels-last activation tensor, fetched at row
// coordinate t0 + (tap - center) * dil: rows outside [0, Lin) (negative included) and channels beyond Cin
// are zero-filled by the TMA unit, which IS the convolution's zero padding - no im2col, no halo code.
// B tiles are boxes {64 ch, BN, 1} of W[g*taps + tap][n][c].
//
// CTA = 192 threads: warp 0 = TMA producer (1 lane), warp 1 = TMEM allocator + MMA issuer (1 lane),
// warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4). One output tile per CTA; shared memory is sized so
// two CTAs co-reside per SM, which overlaps one CTA's epilogue with the other's main loop.
#include "rowgemm_tc.cuh"

#include <mutex>

#include "tc_ptx.cuh"

namespace b200tts {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int MAX_STAGES = 8;

using namespace tc;

struct TcArgs {
  int Cin, N, taps, dil, center, groups, M;
  int BN, stages, kchunks;      // kchunks = ceil(Cin / 64)
  void* out; long o_bstride; int ldo; long o_shift; long o_limit; int out_bf16;
  const float* bias; const float* gate; const float* res; int accumulate; float scale; int act;
  const float* rope_cos; const float* rope_sin; int rope_cols, rope_rows;
  __nv_bfloat16* vt_out; int vt_col0, vt_ld, vt_heads;
};

__global__ void __launch_bounds__(NTHREADS) rowgemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b,
                                                              const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms): [A stages][B stages][barriers][tmem ptr]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = a.BN * BK * 2;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + a.stages * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + a.stages * b_stage_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z / a.groups, g = blockIdx.z % a.groups;
  const int t0 = blockIdx.x * BM, n0 = blockIdx.y * a.BN;
  const int num_kb = a.taps * a.kchunks;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)a.BN) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
      int s = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int j = kb / a.kchunks, c0 = (kb - j * a.kchunks) * BK;
        mbar_wait(&empty_bar[s], phase ^ 1);
        mbar_expect_tx(&full_bar[s], stage_bytes);
        tma_load_3d(smem_a + s * A_STAGE_BYTES, &map_a, &full_bar[s], g * a.Cin + c0, t0 + (j - a.center) * a.dil, b);
        tma_load_3d(smem_b + s * b_stage_bytes, &map_b, &full_bar[s], c0, n0, g * a.taps + j);
        if (++s == a.stages) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor: c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1, K-major both, N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], phase);
        tc_fence_after();
        const uint64_t da = make_desc_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
        const uint64_t db = make_desc_sw128(smem_u32(smem_b + s * b_stage_bytes));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: start-address field += 2
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs have read it
        if (++s == a.stages) { s = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);            // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter q = warp % 4 =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const long obase = (long)b * a.o_bstride;
    const long rowflat = (long)t * a.ldo + (long)g * a.N + a.o_shift;
    const bool row_ok = t < a.M;
    for (int cb = 0; cb < a.BN; cb += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, r);
      tmem_ld_wait();
      const int n = n0 + cb;
      if (!row_ok || n >= a.N) continue;
      const int gn = g * a.N + n;
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        const int nn = n + v4 * 4;
        if (nn >= a.N) break;
        const long flat = rowflat + nn;
        if (flat < 0 || flat >= a.o_limit) continue;
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[v4 * 4 + i]);
        if (a.bias) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + gn + v4 * 4));
          v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
        }
        if (a.act != ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = act_apply(v[i], a.act);
        }
        if (a.rope_cos != nullptr) {
          const int tt = t % a.rope_rows;
          if (nn < a.rope_cols) {            // (x0, x1) -> x*cos + (-x1, x0)*sin, tables repeat per 64-wide head
            const int d = nn & 63;
            const float4 cc = __ldg(reinterpret_cast<const float4*>(a.rope_cos + (long)tt * 64 + d));
            const float4 ss = __ldg(reinterpret_cast<const float4*>(a.rope_sin + (long)tt * 64 + d));
            const float x0 = v[0], x1 = v[1], x2 = v[2], x3 = v[3];
            v[0] = x0 * cc.x - x1 * ss.x; v[1] = x1 * cc.y + x0 * ss.y;
            v[2] = x2 * cc.z - x3 * ss.z; v[3] = x3 * cc.w + x2 * ss.w;
          }
          if (a.vt_out != nullptr && nn >= a.vt_col0) {
            const int cv = nn - a.vt_col0;
            const int hh = cv >> 6, d = cv & 63;
            __nv_bfloat16* o = a.vt_out + ((long)((t / a.rope_rows) * a.vt_heads + hh) * 64 + d) * a.vt_ld + tt;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[(long)i * a.vt_ld] = __float2bfloat16_rn(v[i]);
            continue;
          }
        }
        if (a.gate) {
          const float4 gg = __ldg(reinterpret_cast<const float4*>(a.gate + gn + v4 * 4));
          v[0] *= gg.x; v[1] *= gg.y; v[2] *= gg.z; v[3] *= gg.w;
        }
        if (a.res) {
          const float4 rr = *reinterpret_cast<const float4*>(a.res + obase + flat);
          v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
        }
        if (a.out_bf16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + obase + flat;
          if (a.accumulate) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] += __bfloat162float(o[i]);
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
            const float4 rr = *reinterpret_cast<const float4*>(o);
            v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
          }
          *reinterpret_cast<float4*>(o) = make_float4(v[0] * a.scale, v[1] * a.scale, v[2] * a.scale, v[3] * a.scale);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr)
      fail("cuTensorMapEncodeTiled is not available from the CUDA driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

// bf16 3-D map: dims {d0 (contiguous), d1, d2}, strides in elements {ld1, ld2}, box {64, box1, 1}, 128B swizzle
void tc_encode_map(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                   uint64_t ld2, uint32_t box1) {
  B2_CHECK(((uintptr_t)base & 15) == 0, "TMA base must be 16-byte aligned");
  B2_CHECK((ld1 * 2) % 16 == 0 && (ld2 * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld1 * 2, ld2 * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
}

namespace {

int pick_bn(int N) {
  if (N % 128 == 0) return 128;
  if (N < 128) return (int)round_up(N, 16);    // e.g. 96, 48, 24 -> 32 (rows beyond N are TMA zero fill)
  if (N % 96 == 0) return 96;                  // 192
  if (N % 64 == 0) return 64;
  return 128;                                  // ragged tail tile
}

__global__ void cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long n) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + i));
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(out + i) = pk;
  } else {
    for (long k = i; k < n; ++k) out[k] = __float2bfloat16_rn(in[k]);
  }
}

__global__ void uncast_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

__global__ void cast_pad_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long rows, int C, int ldo) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ldo) return;
  const long r = i / ldo;
  const int c = (int)(i - r * ldo);
  out[i] = c < C ? __float2bfloat16_rn(in[r * C + c]) : __float2bfloat16_rn(0.f);
}

}  // namespace

void cast_f32_to_bf16(const float* in, __nv_bfloat16* out, long n, cudaStream_t s) {
  if (n <= 0) return;
  B2_CHECK(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0, "cast alignment");
  cast_kernel<<<ceil_div(ceil_div(n, 4), 256), 256, 0, s>>>(in, out, n);
  B2_LAUNCH_CHECK(); count_launch();
}

void cast_bf16_to_f32(const __nv_bfloat16* in, float* out, long n, cudaStream_t s) {
  if (n <= 0) return;
  uncast_kernel<<<ceil_div(n, 256), 256, 0, s>>>(in, out, n);
  B2_LAUNCH_CHECK(); count_launch();
}

void cast_pad_f32_to_bf16(const float* in, __nv_bfloat16* out, long rows, int C, int ldo, cudaStream_t s) {
  if (rows <= 0) return;
  cast_pad_kernel<<<ceil_div(rows * ldo, 256), 256, 0, s>>>(in, out, rows, C, ldo);
  B2_LAUNCH_CHECK(); count_launch();
}

void tc_weight_from_f32(TcWeight& tw, const float* w_gjnc, int groups, int taps, int N, int Cin, cudaStream_t s) {
  tw.Cin = Cin; tw.N = N; tw.taps = taps; tw.groups = groups;
  tw.ldc = (int)round_up(Cin, 8);
  const long rows = (long)groups * taps * N;
  tw.w.alloc((size_t)rows * tw.ldc);
  cast_pad_f32_to_bf16(w_gjnc, tw.w.p, rows, Cin, tw.ldc, s);
  tw.BN = pick_bn(N);
  tc_encode_map(&tw.map, tw.w.p, (uint64_t)Cin, (uint64_t)N, (uint64_t)groups * taps, (uint64_t)tw.ldc,
             (uint64_t)N * tw.ldc, (uint32_t)tw.BN);
  tw.ready = true;
}

void rowgemm_tc(const RowGemm& p, const TcWeight& w, cudaStream_t stream) {
  B2_CHECK(w.ready, "rowgemm_tc: tensor-core weights not prepared");
  B2_CHECK(p.Cin == w.Cin && p.N == w.N && p.taps == w.taps && p.groups == w.groups, "rowgemm_tc: weight/problem mismatch");
  B2_CHECK(p.N % 4 == 0 && p.ldo % 4 == 0 && p.o_shift % 4 == 0, "rowgemm_tc: output alignment");
  B2_CHECK(p.ldx % 8 == 0 && p.x_bstride % 8 == 0, "rowgemm_tc: A rows must be 16-byte aligned");
  B2_CHECK(p.groups == 1 || p.Cin % BK == 0, "rowgemm_tc: grouped problems need Cin % 64 == 0");
  B2_CHECK(p.M > 0 && p.B > 0, "rowgemm_tc: empty problem");

  CUtensorMap map_a;
  tc_encode_map(&map_a, p.x, (uint64_t)p.groups * p.Cin, (uint64_t)p.Lin, (uint64_t)p.B, (uint64_t)p.ldx,
             (uint64_t)p.x_bstride, (uint32_t)BM);

  TcArgs a;
  a.Cin = p.Cin; a.N = p.N; a.taps = p.taps; a.dil = p.dil; a.center = p.center; a.groups = p.groups; a.M = p.M;
  a.BN = w.BN; a.kchunks = ceil_div(p.Cin, BK);
  a.out = p.out; a.o_bstride = p.o_bstride; a.ldo = p.ldo; a.o_shift = p.o_shift;
  a.o_limit = p.o_limit ? p.o_limit : (long)p.M * p.ldo;
  a.out_bf16 = p.out_bf16;
  a.bias = p.bias; a.gate = p.gate; a.res = p.res; a.accumulate = p.accumulate; a.scale = p.scale; a.act = p.act;
  a.rope_cos = p.rope_cos; a.rope_sin = p.rope_sin; a.rope_cols = p.rope_cols; a.rope_rows = p.rope_rows > 0 ? p.rope_rows : 1;
  a.vt_out = p.vt_out; a.vt_col0 = p.vt_col0; a.vt_ld = p.vt_ld; a.vt_heads = p.vt_heads;
  B2_CHECK(p.rope_cos == nullptr || (p.rope_sin != nullptr && p.rope_cols % 64 == 0 && p.groups == 1 && p.B == 1),
           "rowgemm_tc: malformed rope epilogue");

  const int stage_bytes = A_STAGE_BYTES + w.BN * BK * 2;
  // keep <= ~110 KB so that two CTAs fit one SM (227 KB): the co-resident CTA hides this one's epilogue
  int stages = (110 * 1024 - 1024 - 256) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const int num_kb = p.taps * a.kchunks;
  if (stages > num_kb) stages = num_kb;
  if (stages < 2 && num_kb >= 2) stages = 2;
  a.stages = stages;
  const int smem = stages * stage_bytes + 1024 /*alignment slack*/ + (2 * MAX_STAGES + 1) * 8 + 16;

  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    B2_CUDA(cudaFuncSetAttribute(rowgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  });
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, w.BN), p.B * p.groups);
  B2_CHECK(grid.y <= 65535 && grid.z <= 65535, "rowgemm_tc grid too large");
  rowgemm_tc_kernel<<<grid, NTHREADS, smem, stream>>>(map_a, w.map, a);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts

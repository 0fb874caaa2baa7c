// Device-side pieces shared by the tcgen05 kernels (rowgemm_tc.cu, dit_chain.cu): tile constants, the argument block of
// one GEMM's epilogue, the smem-transposed coalesced epilogue, and the fully unrolled MMA issue helpers. Include inside
// namespace b200tts, within an anonymous namespace of the including translation unit.
#pragma once
#include <cuda_fp8.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b200tts {
namespace {

constexpr int BK = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NTHREADS3 = 352;         // warps 0..7 epilogue, warp 8 A producer, warp 9 MMA issuer, warp 10 B producer
constexpr int WARP_TMA = 8, WARP_MMA = 9, WARP_TMA_B = 10;   // the SMSP arbiter favours the HIGHEST warp id: the two single-thread roles that
                                            // feed the tensor pipe must not lose issue slots to the epilogue warps they share
                                            // an SMSP with (A/B on one box: +13..25 % on every shape)
constexpr int A_BOX_ROWS = 64;
constexpr int MAX_A_STAGES = 8, MAX_B_STAGES = 8;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;      // one 32x32 fp32 block per epilogue warp
constexpr int EPI_BYTES = 8 * EPI_STAGE_BYTES;

using namespace tc;

enum EpiKind : int { EPI_STD = 0, EPI_ROPE = 1 };

// tcgen05 instruction descriptor of kind::f16: fp32 accumulator, K-major A and B of the 16-bit type (format code 1 = bf16,
// 0 = fp16), N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t idesc_f16kind(int M, int N, int f16) {
  const uint32_t fmt = f16 ? 0u : ((1u << 7) | (1u << 10));
  return (1u << 4) | fmt | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcArgs {
  int Cin, N, taps, dil, center, groups, M;
  int BN, kchunks;              // kchunks = ceil(Cin / 64)
  void* out; long o_bstride; int ldo; long o_shift; long o_limit; int out_bf16;
  const float* bias; const float* gate; const float* res; int accumulate; float scale;
  const __half2* rope_cs; int rope_cols, rope_rows;
  __nv_bfloat16* vt_out; int vt_col0, vt_ld, vt_heads;
  __nv_bfloat16* out2;          // optional 16-bit copy of the output (same indexing, operand dtype)
  const int2* rowinfo;          // ragged batches: (sequence, position) of every output row; null = row / rope_rows, row % rope_rows
  int f16;                      // 16-bit type of the A / B operands, out2 and vt_out: 0 = bf16, 1 = fp16 (IEEE half);
                                // out_bf16 is the type code of `out`: 0 = fp32, 1 = bf16, 2 = fp16
};

struct Tc3Sched {
  int bm;                               // output rows per tile = 128 * halves
  int halves;                           // 128-row accumulators per tile (1, 2 or 4): independent MMA chains
  int m_tiles, n_tiles, num_tiles;      // per (batch, group): m_tiles x n_tiles ; num_tiles = all
  int a_rows;                           // rows per A stage (multiple of 64) = round_up(bm + (taps-1)*dil, 64)
  int a_box;                            // rows per A TMA box: the largest of 256 / 128 / 64 that divides a_rows
  int nA, nB;                           // ring depths (bres: nB = kchunks*taps resident B tiles)
  int bres;                             // 1: the whole weight tensor stays in shared memory for the life of the CTA
  int half_stride, nacc;                // TMEM columns per 128-row accumulator; accumulator stages (1 or 2)
};

__device__ __forceinline__ uint2 pack_bf16x4(float x, float y, float z, float w) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(x, y), p1 = __floats2bfloat162_rn(z, w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  return pk;
}

// 16-bit packing by type code (warp-uniform branch): half != 0 -> IEEE fp16, else bf16
__device__ __forceinline__ uint32_t pack16x2(float x, float y, int half) {
  if (half) { __half2 h = __floats2half2_rn(x, y); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 p = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ uint2 pack16x4(float x, float y, float z, float w, int half) {
  return make_uint2(pack16x2(x, y, half), pack16x2(z, w, half));
}
__device__ __forceinline__ uint16_t pack16(float x, int half) {
  if (half) return __half_as_ushort(__float2half_rn(x));
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float2 unpack16x2(uint32_t v, int half) {
  if (half) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

// Explicit shared-space accesses for the epilogue's staging block. Through a generic pointer the compiler emitted LD.E / ST.E,
// which it may not move across the global stores of the previous row (possible aliasing): the eight staging loads of a block
// were issued one per row, each exposed (ncu r01u: the thin convolutions and the batched DiT GEMMs are epilogue-bound).
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Activations of the bf16 engine (operands are already rounded to 8 mantissa bits, so the 2^-11 MUFU error is noise).
template <int ACT>
__device__ __forceinline__ float act_fast(float v) {
  if (ACT == ACT_GELU_TANH) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float u = k0 * (v + k1 * v * v * v);
    return 0.5f * v * (1.0f + tanh_fast(u));
  } else if (ACT == ACT_MISH) {            // x * tanh(softplus(x)), softplus threshold 20 (F5 modules.py:172)
    const float sp = v > 20.0f ? v : __logf(1.0f + __expf(v));
    return v * tanh_fast(sp);
  } else if (ACT == ACT_GELU_ERF) {
    return 0.5f * v * (1.0f + erff(v * 0.7071067811865476f));
  }
  return v;
}

// One epilogue warp, one accumulator (128 TMEM lanes x BN columns): this warp owns lanes [32q, 32q+32) = tile rows
// row0 .. row0+31 and walks the 32-column blocks cb = cb_first, cb_first + cb_step, ... < BN.
//   phase 1  tcgen05.ld 32 columns of the lane's row -> 8 x STS.128 into the warp's swizzled 4 KB staging block
//   phase 2  lane = (sub = lane/8, c4 = lane%8): rows i*4 + sub (i < 8), columns c4*4 .. c4*4+3 : LDS.128, math, global I/O
struct EpiPos {
  int sub, c4, t_row0, n0;
  long obase, gshift;
};

__device__ __forceinline__ long epi_flat(const TcArgs& a, const EpiPos& p, int i, int n) {
  return (long)(p.t_row0 + i * 4 + p.sub) * a.ldo + p.gshift + n;
}
__device__ __forceinline__ bool epi_ok(const TcArgs& a, const EpiPos& p, int i, int n, long flat) {
  return (p.t_row0 + i * 4 + p.sub) < a.M && n < a.N && flat >= 0 && flat < a.o_limit;
}
// residual operand of block cb, in the phase-2 layout (issued one block ahead: the first before the accumulator is ready,
// the next ones at the end of the previous block's phase 2)
__device__ __forceinline__ void epi_load_res(const TcArgs& a, const EpiPos& p, int cb, float4 (&res)[8]) {
  const int n = p.n0 + cb + p.c4 * 4;
  const bool in = cb < a.BN;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long flat = epi_flat(a, p, i, n);
    res[i] = (in && epi_ok(a, p, i, n, flat)) ? *reinterpret_cast<const float4*>(a.res + p.obase + flat) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// STATS: psum[i] / psq[i] accumulate the sum and the sum of squares of the values this lane writes for its phase-2 row i
// (row i*4 + sub of the warp's 32): the LayerNorm statistics of the fused DiT chain (dit_chain.cu).
template <int KIND, int ACT, bool STATS = false>
__device__ __forceinline__ void epilogue_warp(const TcArgs& a, EpiPos p, uint32_t taddr, uint32_t taddr_hstep, int n_mine, float* stg,
                                              int lane, int g, int cb_first, int cb_step, float4 (&res)[8],
                                              float* psum = nullptr, float* psq = nullptr) {
  // n_mine 128-row halves belong to this warp (rows +256 and TMEM columns +taddr_hstep apart); their 32-column blocks form
  // ONE sequence for the residual look-ahead, so the first block of the second half is prefetched like any other.
  const int sub = p.sub, c4 = p.c4, n0 = p.n0;
  const bool has_res = KIND == EPI_STD && a.res != nullptr;
#pragma unroll 1
  for (int hh = 0; hh < n_mine; ++hh, p.t_row0 += 256, taddr += taddr_hstep) {
  if (p.t_row0 >= a.M) break;                                // warp-uniform: no valid rows in this 32-row block
#pragma unroll 1
  for (int cb = cb_first; cb < a.BN; cb += cb_step) {
    if (n0 + cb >= a.N) break;
    uint32_t r[32];
    tmem_ld32(taddr + (uint32_t)cb, r);
    // this lane's four phase-2 columns: bias / gate fetched under the TMEM load (they are needed first thing in phase 2)
    const int n = n0 + cb + c4 * 4;
    const bool n_ok = n < a.N;
    const float4 bias = (a.bias && n_ok) ? __ldg(reinterpret_cast<const float4*>(a.bias + (long)g * a.N + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 gate = make_float4(1.f, 1.f, 1.f, 1.f);
    if (KIND == EPI_STD && a.gate && n_ok) gate = __ldg(reinterpret_cast<const float4*>(a.gate + (long)g * a.N + n));
    // RoPE block: the (cos, sin) pairs of this lane's 8 phase-2 rows x 4 columns, one 16-byte load each, issued before the
    // TMEM wait (they were 16 dependent L2 round trips inside phase 2: +12 us on the q|k|v GEMM, ncu r01d)
    uint4 cs[8];
    if (KIND == EPI_ROPE && n0 + cb < a.rope_cols) {
      const int d = (n0 + cb + c4 * 4) & 63;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = p.t_row0 + i * 4 + sub;
        const int pos = t < a.M ? (a.rowinfo ? __ldg(&a.rowinfo[t]).y : t % a.rope_rows) : 0;
        cs[i] = t < a.M ? __ldg(reinterpret_cast<const uint4*>(a.rope_cs + (long)pos * 64 + d)) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    tmem_ld_wait();

    if (KIND == EPI_ROPE && a.vt_out != nullptr && n0 + cb >= a.vt_col0) {
      // V columns: written transposed, vt[(batch*heads + h)*64 + d][t]; in the row-per-lane layout consecutive lanes are
      // consecutive t, so each store instruction writes one 64-byte run.
      const int t = p.t_row0 + lane;
      if (t < a.M) {
        int tt = t % a.rope_rows, bb = t / a.rope_rows;
        if (a.rowinfo) { const int2 ri = __ldg(&a.rowinfo[t]); bb = ri.x; tt = ri.y; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = n0 + cb + k * 4;
          if (n < a.N) {
            const float4 bi = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int cv = n - a.vt_col0;
            __nv_bfloat16* o = a.vt_out + ((long)(bb * a.vt_heads + (cv >> 6)) * 64 + (cv & 63)) * a.vt_ld + tt;
            uint16_t* o16 = reinterpret_cast<uint16_t*>(o);
            o16[0] = pack16(__uint_as_float(r[k * 4 + 0]) + bi.x, a.f16);
            o16[(long)a.vt_ld] = pack16(__uint_as_float(r[k * 4 + 1]) + bi.y, a.f16);
            o16[2L * a.vt_ld] = pack16(__uint_as_float(r[k * 4 + 2]) + bi.z, a.f16);
            o16[3L * a.vt_ld] = pack16(__uint_as_float(r[k * 4 + 3]) + bi.w, a.f16);
          }
        }
      }
      continue;
    }

    // phase 1: row `lane`, 16-byte chunk k goes to chunk slot k ^ (lane & 7) (conflict-free per quarter warp)
    const uint32_t stg_s = smem_u32(stg);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      sts128(stg_s + (uint32_t)(lane * 32 + ((k ^ (lane & 7)) << 2)) * 4u, r[k * 4], r[k * 4 + 1], r[k * 4 + 2], r[k * 4 + 3]);
    __syncwarp();

    // the next block's residual: issued here, once r[] is dead, so it has the whole of phase 2 plus the next block's TMEM
    // load and transpose to arrive (issued after phase 2 it was exposed: the thin BigVGAN conv2 ran at half speed)
    float4 res_n[8];
    if (has_res) {
      const bool last_cb = cb + cb_step >= a.BN || n0 + cb + cb_step >= a.N;
      EpiPos pn = p;
      if (last_cb) pn.t_row0 += 256;                         // (rows beyond M / a half that is not ours load zeros)
      epi_load_res(a, pn, (last_cb && hh + 1 < n_mine) ? cb_first : cb + cb_step, res_n);
    }

    // phase 2
    const bool rope = KIND == EPI_ROPE && n < a.rope_cols;
    float4 accs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                 // all eight staging rows of this lane in flight at once
      const int row = i * 4 + sub;
      accs[i] = lds128(stg_s + (uint32_t)(row * 32 + ((c4 ^ (row & 7)) << 2)) * 4u);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long flat = epi_flat(a, p, i, n);
      const bool ok = epi_ok(a, p, i, n, flat);
      const float4 acc = accs[i];
      float v0 = acc.x + bias.x, v1 = acc.y + bias.y, v2 = acc.z + bias.z, v3 = acc.w + bias.w;
      if (KIND == EPI_ROPE) {
        if (rope) {                             // (x0, x1) -> x*cos + (-x1, x0)*sin, tables repeat per 64-wide head
          const float2 cs0 = __half22float2(*reinterpret_cast<const __half2*>(&cs[i].x));
          const float2 cs1 = __half22float2(*reinterpret_cast<const __half2*>(&cs[i].y));
          const float2 cs2 = __half22float2(*reinterpret_cast<const __half2*>(&cs[i].z));
          const float2 cs3 = __half22float2(*reinterpret_cast<const __half2*>(&cs[i].w));
          const float x0 = v0, x1 = v1, x2 = v2, x3 = v3;
          v0 = x0 * cs0.x - x1 * cs0.y; v1 = x1 * cs1.x + x0 * cs1.y;
          v2 = x2 * cs2.x - x3 * cs2.y; v3 = x3 * cs3.x + x2 * cs3.y;
        }
      } else {
        if (ACT != ACT_NONE) { v0 = act_fast<ACT>(v0); v1 = act_fast<ACT>(v1); v2 = act_fast<ACT>(v2); v3 = act_fast<ACT>(v3); }
        v0 *= gate.x; v1 *= gate.y; v2 *= gate.z; v3 *= gate.w;
        if (has_res) { v0 += res[i].x; v1 += res[i].y; v2 += res[i].z; v3 += res[i].w; }
      }
      if (STATS && ok) {
        const float w0 = v0 * a.scale, w1 = v1 * a.scale, w2 = v2 * a.scale, w3 = v3 * a.scale;
        psum[i] += (w0 + w1) + (w2 + w3);
        psq[i] += (w0 * w0 + w1 * w1) + (w2 * w2 + w3 * w3);
      }
      if (ok) {
        if (a.out_bf16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + p.obase + flat;
          if (KIND == EPI_STD && a.accumulate) {
            const uint2 pv = *reinterpret_cast<const uint2*>(o);
            const float2 p01 = unpack16x2(pv.x, a.out_bf16 == 2), p23 = unpack16x2(pv.y, a.out_bf16 == 2);
            v0 += p01.x; v1 += p01.y; v2 += p23.x; v3 += p23.y;
          }
          *reinterpret_cast<uint2*>(o) = pack16x4(v0 * a.scale, v1 * a.scale, v2 * a.scale, v3 * a.scale, a.out_bf16 == 2);
        } else {
          float* o = reinterpret_cast<float*>(a.out) + p.obase + flat;
          if (KIND == EPI_STD && a.accumulate) {
            const float4 pv = *reinterpret_cast<const float4*>(o);
            v0 += pv.x; v1 += pv.y; v2 += pv.z; v3 += pv.w;
          }
          *reinterpret_cast<float4*>(o) = make_float4(v0 * a.scale, v1 * a.scale, v2 * a.scale, v3 * a.scale);
        }
        if (KIND == EPI_STD && a.out2 != nullptr)      // second copy of the result in bf16 (the next GEMM's A operand)
          *reinterpret_cast<uint2*>(a.out2 + p.obase + flat) = pack16x4(v0 * a.scale, v1 * a.scale, v2 * a.scale, v3 * a.scale, a.f16);
      }
    }
    if (has_res) {
#pragma unroll
      for (int i = 0; i < 8; ++i) res[i] = res_n[i];
    }
    __syncwarp();                                // the staging block is rewritten by the next iteration's phase 1
  }
  }
}

// =============================================================================================
// Row-layout epilogue with TMA tiles (dit_chain.cu). The transposing epilogue above spends ~600 instructions per 32x32
// block on address arithmetic, bounds tests, the shared-memory transpose and 16 + 8 global accesses per lane -- 2.3 us per
// block when the tile's epilogue is exposed (one tile per CTA: chain timeline, profiles/r02). Here a lane keeps the
// accumulator's native layout (lane = row, 32 consecutive columns in registers): bias / gate come from a shared-memory copy
// (broadcast reads), the fp32 residual tile is fetched by the TMA unit into a 128B-swizzled tile (8 conflict-free LDS.128
// per lane), the result goes into a swizzled staging tile (8 or 4 STS.128) and ONE thread hands it to the TMA unit, which does
// the coalescing and the row clipping. LayerNorm statistics are thread-local sums (a lane owns a row): no shuffles.
// =============================================================================================
struct EpiTile {
  uint8_t* res_tile;      // 4 KB, 1024-byte aligned: residual tile [32 rows][32 fp32], SWIZZLE_128B
  uint8_t* out_tile;      // 4 KB, 1024-byte aligned: result tile, fp32 [32][128 B] SWIZZLE_128B or 16-bit [32][64 B] SWIZZLE_64B
  uint64_t* res_bar;      // mbarrier (count 1) the residual tiles complete on
  uint32_t res_phase;     // parity of the next residual tile to wait for
};
enum TileKind : int { TK_RES_F32 = 0, TK_ACT16 = 1, TK_ROPE16 = 2 };

// lane 0: fetch the residual tile of the 32-column block at global column n, rows row0 .. row0 + 31
__device__ __forceinline__ void epi_tma_fetch_res(const CUtensorMap* map_res, EpiTile& et, int n, int row0, int lane) {
  if (lane == 0) {
    mbar_expect_tx(et.res_bar, 32u * 128u);
    tma_load_2d(et.res_tile, map_res, et.res_bar, n, row0);
  }
}
// the warp's stores have been performed (global memory holds the tiles); orders them before generic-proxy reads that follow a
// barrier (LN pass 2) and before the team counters
__device__ __forceinline__ void epi_tma_drain(int lane) {
  if (lane == 0) {
    bulk_wait0();
    asm volatile("fence.proxy.async.global;" ::: "memory");
  }
  __syncwarp();
}

// One warp, rows row0 .. row0 + 31 (TMEM lanes of `taddr`), blocks cb = cb_first, cb_first + cb_step, ... < a.BN of the CTA's
// column slice starting at global column n0. s_bias / s_gate: shared-memory copies of bias[n0 ..), gate[n0 ..) (s_gate may be
// null). TK_RES_F32: out = res + gate * (acc + bias) as fp32 (the residual tile of the FIRST block must already be in flight:
// epi_tma_fetch_res); TK_ACT16: out = act(acc + bias) as 16 bit; TK_ROPE16: q | k columns roped as 16 bit, V columns through
// the transposed scalar path of epilogue_warp.
// AFFINE (the LayerNorm folded into the GEMM, dit_chain.cu): the A operand held c * x * (1 + scale) instead of the normalised
// rows, so  LN(x)(1 + scale) + shift) W^T + b  =  rho * acc - rmu * u[col] + v[col]  with rho = rstd / c, rmu = rstd * mean and
// the per-column vectors u = W (1 + scale), v = W shift + b passed in place of gate / bias (s_gate = u, s_bias = v).
// EpiEmit (TK_RES_F32 only): besides the fp32 result x, write c * x * mul[col] as 16 bit straight to global memory (a lane
// owns 64 contiguous bytes of its row) -- the A operand of the NEXT GEMM, produced without a second pass over x.
struct EpiEmit {
  uint16_t* row = nullptr;     // this lane's row of the 16-bit tensor (null: row out of range / nothing to emit)
  uint8_t* row8 = nullptr;     // ... or of the 8-bit (e4m3) tensor: the next GEMM runs with fp8 operands
  const float* s_mul = nullptr;   // shared-memory copy of (1 + scale)[n0 ..)
                                  // (AFFINE epilogues reuse the field: per-column multiplier of the accumulator = e4m3 weight scales)
  float c = 1.0f;              // the row's scale (a stale 1 / std estimate: keeps fp16 / e4m3 in range)
};
template <int TK, int ACT, bool STATS, bool AFFINE = false, bool FP8 = false>
__device__ __forceinline__ void epilogue_rows_tma(const TcArgs& a, const CUtensorMap* map_out, const CUtensorMap* map_res, EpiTile& et,
                                                  uint32_t taddr, int row0, int n0, int cb_first, int cb_step, int lane,
                                                  const float* s_bias, const float* s_gate, float& rsum, float& rsq,
                                                  uint64_t store_policy = 0ull /* L2 cache hint of the result tiles, 0 = none */,
                                                  float* stat_slots = nullptr /* STATS: [group][parity][128 rows] float2 in smem */,
                                                  int stat_col0 = 0 /* first column of the caller's slice */, int stat_row = 0,
                                                  float rho = 1.0f, float rmu = 0.0f, EpiEmit emit = EpiEmit()) {
  const uint32_t out_s = smem_u32(et.out_tile), res_s = smem_u32(et.res_tile);
  const uint32_t bias_s = smem_u32(s_bias), gate_s = s_gate ? smem_u32(s_gate) : 0u;
  const uint32_t row128 = (uint32_t)lane * 128u, sw128 = (uint32_t)(lane & 7);
  const uint32_t row64 = (uint32_t)lane * 64u, sw64 = (uint32_t)((lane >> 1) & 3);
  const int t = row0 + lane;
  int tt = 0, bb = 0;                                          // position in / index of this row's sequence (RoPE, V^T)
  if (TK == TK_ROPE16 && t < a.M) {
    if (a.rowinfo) { const int2 ri = __ldg(&a.rowinfo[t]); bb = ri.x; tt = ri.y; }
    else { tt = t % a.rope_rows; bb = t / a.rope_rows; }
  }
#pragma unroll 1
  for (int cb = cb_first; cb < a.BN; cb += cb_step) {
    const int n = n0 + cb;
    if (n >= a.N) break;
    uint32_t v[32];
    tmem_ld32(taddr + (uint32_t)cb, v);
    uint4 cs[8];
    const bool rope = TK == TK_ROPE16 && n < a.rope_cols;
    if (rope) {           // this row's (cos, sin) pairs of the block's 32 columns: one 128-byte line, fetched under the TMEM load
      const __half2* cp = a.rope_cs + (long)tt * 64 + (n & 63);
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] = __ldg(reinterpret_cast<const uint4*>(cp) + k);
    }
    if (TK == TK_RES_F32) { mbar_wait(et.res_bar, et.res_phase); et.res_phase ^= 1u; }
    tmem_ld_wait();

    if (TK == TK_ROPE16 && a.vt_out != nullptr && n >= a.vt_col0) {
      // V columns: written transposed, vt[(sequence*heads + h)*64 + d][t]; consecutive lanes are consecutive t
      if (t < a.M) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 bi = lds128(bias_s + (uint32_t)(cb + k * 4) * 4u);
          float w0 = __uint_as_float(v[k * 4 + 0]), w1 = __uint_as_float(v[k * 4 + 1]), w2 = __uint_as_float(v[k * 4 + 2]), w3 = __uint_as_float(v[k * 4 + 3]);
          if (AFFINE) {
            const float4 u = lds128(gate_s + (uint32_t)(cb + k * 4) * 4u);
            bi.x = fmaf(-rmu, u.x, bi.x); bi.y = fmaf(-rmu, u.y, bi.y); bi.z = fmaf(-rmu, u.z, bi.z); bi.w = fmaf(-rmu, u.w, bi.w);
            float4 r4 = make_float4(rho, rho, rho, rho);
            if (FP8 && emit.s_mul != nullptr) {
              const float4 sw = lds128(smem_u32(emit.s_mul) + (uint32_t)(cb + k * 4) * 4u);
              r4.x *= sw.x; r4.y *= sw.y; r4.z *= sw.z; r4.w *= sw.w;
            }
            w0 *= r4.x; w1 *= r4.y; w2 *= r4.z; w3 *= r4.w;
          }
          const int cv = n + k * 4 - a.vt_col0;
          uint16_t* o16 = reinterpret_cast<uint16_t*>(a.vt_out) + ((long)(bb * a.vt_heads + (cv >> 6)) * 64 + (cv & 63)) * a.vt_ld + tt;
          o16[0] = pack16(w0 + bi.x, a.f16);
          o16[(long)a.vt_ld] = pack16(w1 + bi.y, a.f16);
          o16[2L * a.vt_ld] = pack16(w2 + bi.z, a.f16);
          o16[3L * a.vt_ld] = pack16(w3 + bi.w, a.f16);
        }
      }
      continue;
    }

    float f[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 bi = lds128(bias_s + (uint32_t)(cb + k * 4) * 4u);
      if (AFFINE) {
        const float4 u = lds128(gate_s + (uint32_t)(cb + k * 4) * 4u);
        float4 r4 = make_float4(rho, rho, rho, rho);
        if (FP8 && emit.s_mul != nullptr) {                      // e4m3 weights: their per-output-channel scales
          const float4 sw = lds128(smem_u32(emit.s_mul) + (uint32_t)(cb + k * 4) * 4u);
          r4.x *= sw.x; r4.y *= sw.y; r4.z *= sw.z; r4.w *= sw.w;
        }
        f[k * 4 + 0] = fmaf(r4.x, __uint_as_float(v[k * 4 + 0]), fmaf(-rmu, u.x, bi.x));
        f[k * 4 + 1] = fmaf(r4.y, __uint_as_float(v[k * 4 + 1]), fmaf(-rmu, u.y, bi.y));
        f[k * 4 + 2] = fmaf(r4.z, __uint_as_float(v[k * 4 + 2]), fmaf(-rmu, u.z, bi.z));
        f[k * 4 + 3] = fmaf(r4.w, __uint_as_float(v[k * 4 + 3]), fmaf(-rmu, u.w, bi.w));
      } else {
        f[k * 4 + 0] = __uint_as_float(v[k * 4 + 0]) + bi.x; f[k * 4 + 1] = __uint_as_float(v[k * 4 + 1]) + bi.y;
        f[k * 4 + 2] = __uint_as_float(v[k * 4 + 2]) + bi.z; f[k * 4 + 3] = __uint_as_float(v[k * 4 + 3]) + bi.w;
      }
    }
    if (TK == TK_ACT16 && ACT != ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = act_fast<ACT>(f[i]);
    }
    if (TK == TK_RES_F32) {
      if (s_gate) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 g = lds128(gate_s + (uint32_t)(cb + k * 4) * 4u);
          f[k * 4 + 0] *= g.x; f[k * 4 + 1] *= g.y; f[k * 4 + 2] *= g.z; f[k * 4 + 3] *= g.w;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 rr = lds128(res_s + row128 + (((uint32_t)k ^ sw128) << 4));
        f[k * 4 + 0] += rr.x; f[k * 4 + 1] += rr.y; f[k * 4 + 2] += rr.z; f[k * 4 + 3] += rr.w;
      }
      __syncwarp();                                            // every lane has read the residual tile: fetch the next one
      if (cb + cb_step < a.BN && n + cb_step < a.N) epi_tma_fetch_res(map_res, et, n + cb_step, row0, lane);
    }
    if (rope) {           // (x0, x1) -> x * cos + (-x1, x0) * sin, interleaved pairs (F5 modules.py:421-430)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t w[4] = {cs[k].x, cs[k].y, cs[k].z, cs[k].w};
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const float2 c0 = __half22float2(*reinterpret_cast<const __half2*>(&w[2 * h2]));
          const float2 c1 = __half22float2(*reinterpret_cast<const __half2*>(&w[2 * h2 + 1]));
          const float x0 = f[k * 4 + 2 * h2], x1 = f[k * 4 + 2 * h2 + 1];
          f[k * 4 + 2 * h2] = x0 * c0.x - x1 * c0.y;
          f[k * 4 + 2 * h2 + 1] = x1 * c1.x + x0 * c1.y;
        }
      }
    }
    if (STATS) {
      // LayerNorm partials of this row, one (sum, sum of squares) per 128-column group of the OUTPUT and column parity of the
      // warp: a group is two of this warp's blocks (cb & 64 = 0, then != 0), summed in one chain from zero and parked in shared
      // memory. The grouping is the same whatever the slice width of the caller, so the statistics of a row do not depend on
      // how many CTA pairs share its row block (batch == single utterance, bit for bit).
#pragma unroll
      for (int i = 0; i < 32; ++i) { rsum += f[i]; rsq = fmaf(f[i], f[i], rsq); }
      if ((cb & 64) != 0 && stat_slots != nullptr) {
        float2* slot = reinterpret_cast<float2*>(stat_slots) + (size_t)(((n - stat_col0) >> 7) * 2 + ((cb >> 5) & 1)) * 128 + stat_row;
        *slot = make_float2(rsum, rsq);
        rsum = 0.f; rsq = 0.f;
      }
    }
    if (FP8 && TK == TK_RES_F32 && emit.row8 != nullptr) {
      // e4m3 operand of the next GEMM: 32 columns = 32 bytes per lane; cvt saturates at +-448
      const uint32_t mul_s = smem_u32(emit.s_mul);
      uint32_t w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 m0 = lds128(mul_s + (uint32_t)(cb + k * 4) * 4u);
        const float y0 = __fmul_rn(__fmul_rn(emit.c, f[k * 4 + 0]), m0.x), y1 = __fmul_rn(__fmul_rn(emit.c, f[k * 4 + 1]), m0.y);
        const float y2 = __fmul_rn(__fmul_rn(emit.c, f[k * 4 + 2]), m0.z), y3 = __fmul_rn(__fmul_rn(emit.c, f[k * 4 + 3]), m0.w);
        const uint32_t lo = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(y0, y1), __NV_SATFINITE, __NV_E4M3);
        const uint32_t hi = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(y2, y3), __NV_SATFINITE, __NV_E4M3);
        w[k] = lo | (hi << 16);
      }
      *reinterpret_cast<uint4*>(emit.row8 + n) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(emit.row8 + n + 16) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    if (TK == TK_RES_F32 && emit.row != nullptr) {
      const uint32_t mul_s = smem_u32(emit.s_mul);
      const float lim = a.f16 ? 65504.0f : 3.0e38f;            // saturate instead of overflowing to inf (fp16 operands)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 m0 = lds128(mul_s + (uint32_t)(cb + k * 8) * 4u), m1 = lds128(mul_s + (uint32_t)(cb + k * 8 + 4) * 4u);
        const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float y0 = fminf(fmaxf(__fmul_rn(__fmul_rn(emit.c, f[k * 8 + 2 * i]), mm[2 * i]), -lim), lim);
          const float y1 = fminf(fmaxf(__fmul_rn(__fmul_rn(emit.c, f[k * 8 + 2 * i + 1]), mm[2 * i + 1]), -lim), lim);
          w[i] = pack16x2(y0, y1, a.f16);
        }
        *reinterpret_cast<uint4*>(emit.row + n + k * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (FP8 && TK == TK_ACT16 && emit.c != 0.0f) {
      // e4m3 hidden activation (the A operand of an e4m3 ff2): emit.c * act(.), 32 bytes per lane straight to global memory
      // (emit.row8 = this lane's row, null beyond the last row); no 16-bit tile is written
      if (emit.row8 != nullptr) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t lo = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(emit.c * f[k * 4 + 0], emit.c * f[k * 4 + 1]), __NV_SATFINITE, __NV_E4M3);
          const uint32_t hi = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(emit.c * f[k * 4 + 2], emit.c * f[k * 4 + 3]), __NV_SATFINITE, __NV_E4M3);
          w[k] = lo | (hi << 16);
        }
        *reinterpret_cast<uint4*>(emit.row8 + n) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(emit.row8 + n + 16) = make_uint4(w[4], w[5], w[6], w[7]);
      }
      continue;
    }
    if (lane == 0) bulk_wait_read0();                          // the previous block's store has read the staging tile
    __syncwarp();
    if (TK == TK_RES_F32) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sts128(out_s + row128 + (((uint32_t)k ^ sw128) << 4), __float_as_uint(f[k * 4]), __float_as_uint(f[k * 4 + 1]),
               __float_as_uint(f[k * 4 + 2]), __float_as_uint(f[k * 4 + 3]));
    } else {
      const int half = a.f16;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sts128(out_s + row64 + (((uint32_t)k ^ sw64) << 4), pack16x2(f[k * 8], f[k * 8 + 1], half), pack16x2(f[k * 8 + 2], f[k * 8 + 3], half),
               pack16x2(f[k * 8 + 4], f[k * 8 + 5], half), pack16x2(f[k * 8 + 6], f[k * 8 + 7], half));
    }
    fence_proxy_async();                                       // generic-proxy writes of the staging tile -> async proxy
    __syncwarp();
    if (lane == 0) {
      if (store_policy) tma_store_2d_hint(map_out, et.out_tile, n, row0, store_policy);
      else tma_store_2d(map_out, et.out_tile, n, row0);
      bulk_commit();
    }
  }
}

// The KS x NH MMAs of one (64-channel chunk, tap): K steps outermost, 128-row halves innermost. KS = 4 for a full chunk,
// 2 / 3 for the ragged last chunk of the thin BigVGAN stages (C = 24, 48, 96), where the generic loop was the critical path.
template <int NH, int KS>
__device__ __forceinline__ void issue_tap(uint32_t d, uint32_t hstep, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
#pragma unroll
  for (int k = 0; k < KS; ++k)
#pragma unroll
    for (int h = 0; h < NH; ++h)
      umma_bf16_lohi(d + (uint32_t)h * hstep, a_lo + (uint32_t)h * 1024u + 2u * k, b_lo + 2u * k, idesc, k == 0 ? accum : 1u);
}
template <int NH>
__device__ __forceinline__ bool issue_tap_ks(int ksteps, uint32_t d, uint32_t hstep, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accum) {
  if (ksteps == 4) issue_tap<NH, 4>(d, hstep, a_lo, b_lo, idesc, accum);
  else if (ksteps == 2) issue_tap<NH, 2>(d, hstep, a_lo, b_lo, idesc, accum);
  else if (ksteps == 3) issue_tap<NH, 3>(d, hstep, a_lo, b_lo, idesc, accum);
  else return false;
  return true;
}

}  // namespace
}  // namespace b200tts

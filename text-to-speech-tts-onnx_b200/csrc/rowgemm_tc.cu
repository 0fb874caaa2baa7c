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
// Two kernels share the epilogue:
//  * rowgemm_tc2_kernel (default): persistent, one CTA per SM, 256 x BN output tiles (two M=128 accumulators that
//    share every B tile), double-buffered accumulators in TMEM (512 columns) so the epilogue of tile i overlaps the
//    main loop of tile i+1. The A operand of a convolution is fetched ONCE per 64-channel chunk as a
//    (256 + (taps-1)*dil)-row halo tile; every tap is an MMA whose A descriptor starts (tap*dil) rows further down
//    that tile (plain descriptor start-address shift; the swizzle XOR is a function of the absolute smem address), so L2->SMEM traffic per FLOP
//    is ~1/256 + 1/(taps*BN) B instead of v1's 1/128 + 1/128 (which made v1 L2-bandwidth-bound at ~25 % of peak).
//    320 threads: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = epilogue.
//  * rowgemm_tc_kernel (v1, B200TTS_GEMM=v1): one 128 x BN tile per 192-thread CTA, per-tap A tiles.
#include "rowgemm_tc.cuh"

#include <cstdlib>
#include <mutex>
#include <string>

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
  __nv_bfloat16* out2;          // optional bf16 copy of the output (same indexing)
};

// Epilogue of 16 consecutive output columns [n, n+16) of one output row t (registers r[] straight from tcgen05.ld):
// bias, activation, fused q/k RoPE + transposed V store, gate, residual, accumulate, scale, fp32 or bf16 store.
__device__ __forceinline__ void epilogue_chunk16(const TcArgs& a, const uint32_t (&r)[16], int b, int g, int t, int n,
                                                 long obase, long rowflat) {
  if (n >= a.N) return;
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
    if (a.out2 != nullptr) {             // second copy of the result in bf16 (the next GEMM's A operand)
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0] * a.scale, v[1] * a.scale);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2] * a.scale, v[3] * a.scale);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0);
      pk.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(a.out2 + obase + flat) = pk;
    }
  }
}

// ---- v2 epilogue: 32 output columns [n, n+32) of one row per call, loads batched ahead of the math -------------------
// The first version walked 4-column groups with bias -> gate -> residual loads each waiting on the previous group's
// store (possible aliasing), which serialised ~3 L2 round trips per group: 70k cycles per 256x128 tile on the DiT
// GEMMs (ncu source page, profiles/r01). Now: the residual of the NEXT chunk is prefetched into registers while this
// chunk is processed (and the first chunk's before the accumulator is even complete), bias / gate / RoPE tables are
// fetched through the read-only path in one batch, then math, then stores.
__device__ __forceinline__ void epi_prefetch_res(const TcArgs& a, bool row_ok, long obase, long rowflat, int n, float4 (&res)[8]) {
#pragma unroll
  for (int v4 = 0; v4 < 8; ++v4) {
    const int nn = n + v4 * 4;
    const long flat = rowflat + nn;
    const bool ok = row_ok && nn < a.N && flat >= 0 && flat < a.o_limit;
    res[v4] = ok ? *reinterpret_cast<const float4*>(a.res + obase + flat) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ uint2 pack_bf16x4(float x, float y, float z, float w) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(x, y), p1 = __floats2bfloat162_rn(z, w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  return pk;
}

// ROPE = true: the fused q/k/v epilogue (bias, interleaved RoPE on columns < rope_cols, V written transposed); otherwise the
// general one (bias, activation, gate, residual, accumulate, scale, optional second bf16 copy). Two instantiations keep the
// live register set of each under the 168-register cap of a 320-thread CTA.
template <bool ROPE>
__device__ __forceinline__ void epilogue_chunk32(const TcArgs& a, const uint32_t (&r)[32], const float4 (&res)[8], int g, int t,
                                                 int n, long obase, long rowflat) {
  if (n >= a.N) return;
  const int gn = g * a.N + n;
  const bool rope = ROPE && n < a.rope_cols;
  const int tt = ROPE ? t % a.rope_rows : 0;
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {         // two sub-batches of 16 columns: loads first, then math + stores
    float4 bias[4], gate[4], rc[4], rs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = hb * 16 + u * 4;
      const bool in = n + c < a.N;
      bias[u] = (a.bias && in) ? __ldg(reinterpret_cast<const float4*>(a.bias + gn + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (!ROPE && a.gate) gate[u] = in ? __ldg(reinterpret_cast<const float4*>(a.gate + gn + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (rope) {
        const int d = (n + c) & 63;
        rc[u] = __ldg(reinterpret_cast<const float4*>(a.rope_cos + (long)tt * 64 + d));
        rs[u] = __ldg(reinterpret_cast<const float4*>(a.rope_sin + (long)tt * 64 + d));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v4 = hb * 4 + u;
      const int nn = n + v4 * 4;
      const long flat = rowflat + nn;
      if (nn >= a.N || flat < 0 || flat >= a.o_limit) continue;
      float v[4];
      v[0] = __uint_as_float(r[v4 * 4 + 0]) + bias[u].x; v[1] = __uint_as_float(r[v4 * 4 + 1]) + bias[u].y;
      v[2] = __uint_as_float(r[v4 * 4 + 2]) + bias[u].z; v[3] = __uint_as_float(r[v4 * 4 + 3]) + bias[u].w;
      if (ROPE) {
        if (rope) {                          // (x0, x1) -> x*cos + (-x1, x0)*sin, tables repeat per 64-wide head
          const float x0 = v[0], x1 = v[1], x2 = v[2], x3 = v[3];
          v[0] = x0 * rc[u].x - x1 * rs[u].x; v[1] = x1 * rc[u].y + x0 * rs[u].y;
          v[2] = x2 * rc[u].z - x3 * rs[u].z; v[3] = x3 * rc[u].w + x2 * rs[u].w;
        }
        if (a.vt_out != nullptr && nn >= a.vt_col0) {
          const int cv = nn - a.vt_col0;
          const int hh = cv >> 6, d = cv & 63;
          __nv_bfloat16* o = a.vt_out + ((long)((t / a.rope_rows) * a.vt_heads + hh) * 64 + d) * a.vt_ld + tt;
#pragma unroll
          for (int i = 0; i < 4; ++i) o[(long)i * a.vt_ld] = __float2bfloat16_rn(v[i]);
          continue;
        }
      } else {
        if (a.act != ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = act_apply(v[i], a.act);
        }
        if (a.gate) { v[0] *= gate[u].x; v[1] *= gate[u].y; v[2] *= gate[u].z; v[3] *= gate[u].w; }
        if (a.res) { v[0] += res[v4].x; v[1] += res[v4].y; v[2] += res[v4].z; v[3] += res[v4].w; }
      }
      if (a.out_bf16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + obase + flat;
        if (!ROPE && a.accumulate) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] += __bfloat162float(o[i]);
        }
        *reinterpret_cast<uint2*>(o) = pack_bf16x4(v[0] * a.scale, v[1] * a.scale, v[2] * a.scale, v[3] * a.scale);
      } else {
        float* o = reinterpret_cast<float*>(a.out) + obase + flat;
        if (!ROPE && a.accumulate) {
          const float4 rr = *reinterpret_cast<const float4*>(o);
          v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
        }
        *reinterpret_cast<float4*>(o) = make_float4(v[0] * a.scale, v[1] * a.scale, v[2] * a.scale, v[3] * a.scale);
      }
      if (!ROPE && a.out2 != nullptr)      // second copy of the result in bf16 (the next GEMM's A operand)
        *reinterpret_cast<uint2*>(a.out2 + obase + flat) = pack_bf16x4(v[0] * a.scale, v[1] * a.scale, v[2] * a.scale, v[3] * a.scale);
    }
  }
}

// =============================================================================================
// v2: persistent 256 x BN tiles, halo A tiles, double-buffered TMEM accumulators
// =============================================================================================
constexpr int BM2 = 256;
constexpr int NTHREADS2 = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int A_BOX_ROWS = 64;
constexpr int MAX_A_STAGES = 4, MAX_B_STAGES = 8;

struct Tc2Sched {
  int m_tiles, n_tiles, num_tiles;      // per (batch, group): m_tiles x n_tiles ; num_tiles = all
  int a_rows;                           // rows per A stage (multiple of 64) = round_up(256 + (taps-1)*dil, 64)
  int nA, nB;                           // ring depths
  int half_stride, nacc;                // TMEM columns between the two M halves; accumulator stages (1 or 2)
};

// A tap's A tile starts (tap*dil) rows = (tap*dil)*128 bytes into the halo tile, i.e. generally NOT on a 1024-byte
// swizzle-atom boundary. Measured on B200 (tools/debug_conv.py): the tensor core applies the 128B-swizzle XOR to the
// absolute shared-memory address bits, exactly as the TMA unit did when it wrote the tile, so the plain descriptor with
// the shifted start address is correct and the matrix-base-offset field must stay 0 (setting it to the row phase
// (addr >> 7) & 7 double-applies the rotation and corrupts every tap whose shift is not a multiple of 8 rows).

template <bool ROPE>
__global__ void __launch_bounds__(NTHREADS2, 1) rowgemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_b,
                                                                   const TcArgs a, const Tc2Sched sc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_stage_bytes = sc.a_rows * 128;
  const int b_stage_bytes = a.BN * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + sc.nA * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + sc.nB * b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + MAX_A_STAGES;
  uint64_t* b_full = a_empty + MAX_A_STAGES;
  uint64_t* b_empty = b_full + MAX_B_STAGES;
  uint64_t* acc_full = b_empty + MAX_B_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < sc.nA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < sc.nB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_bg = sc.m_tiles * sc.n_tiles;
  const int halo_lo = a.center * a.dil;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      const int a_boxes = sc.a_rows / A_BOX_ROWS;
      for (int tile = blockIdx.x; tile < sc.num_tiles; tile += gridDim.x) {
        const int bg = tile / per_bg, rem = tile - bg * per_bg;
        const int nt = rem / sc.m_tiles, mt = rem - nt * sc.m_tiles;
        const int b = bg / a.groups, g = bg - b * a.groups;
        const int t0 = mt * BM2, n0 = nt * a.BN;
        for (int c = 0; c < a.kchunks; ++c) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], (uint32_t)a_stage_bytes);
          for (int rbx = 0; rbx < a_boxes; ++rbx)
            tma_load_3d(smem_a + sa * a_stage_bytes + rbx * (A_BOX_ROWS * 128), &map_a, &a_full[sa], g * a.Cin + c * BK,
                        t0 - halo_lo + rbx * A_BOX_ROWS, b);
          if (++sa == sc.nA) { sa = 0; pa ^= 1; }
          for (int j = 0; j < a.taps; ++j) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_expect_tx(&b_full[sb], (uint32_t)b_stage_bytes);
            tma_load_3d(smem_b + sb * b_stage_bytes, &map_b, &b_full[sb], c * BK, n0, g * a.taps + j);
            if (++sb == sc.nB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the (warp-uniform) loops, one elected lane issues =====
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a)), b_lo0 = desc_lo_sw128(smem_u32(smem_b));
    const uint32_t a_stage_lo = (uint32_t)a_stage_bytes >> 4, b_stage_lo = (uint32_t)b_stage_bytes >> 4;
    const uint32_t tap_lo = (uint32_t)a.dil * 8u;                   // dil rows x 128 B, in 16-byte units
    const uint32_t hs = (uint32_t)sc.half_stride;
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < sc.num_tiles; tile += gridDim.x, ++it) {
      const int rem = tile % per_bg;
      const int mt = rem % sc.m_tiles;
      const bool two = (mt * BM2 + 128 < a.M);                      // skip the second M half of a ragged last tile
      const int acc = sc.nacc == 2 ? (it & 1) : 0;
      const uint32_t accphase = sc.nacc == 2 ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      mbar_wait(&acc_empty[acc], accphase ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)acc * 2u * hs, d1 = d0 + hs;
      for (int c = 0; c < a.kchunks; ++c) {
        int ksteps = (a.Cin - c * BK + UMMA_K - 1) / UMMA_K;
        if (ksteps > BK / UMMA_K) ksteps = BK / UMMA_K;
        mbar_wait(&a_full[sa], pa);
        uint32_t a_lo = a_lo0 + (uint32_t)sa * a_stage_lo;
        for (int j = 0; j < a.taps; ++j, a_lo += tap_lo) {
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_stage_lo;
          const uint32_t first = (c > 0 || j > 0) ? 1u : 0u;
          if (elect_one()) {
            if (ksteps == 4) {
              umma_bf16_lohi(d0, a_lo + 0, b_lo + 0, idesc, first);
              umma_bf16_lohi(d0, a_lo + 2, b_lo + 2, idesc, 1u);
              umma_bf16_lohi(d0, a_lo + 4, b_lo + 4, idesc, 1u);
              umma_bf16_lohi(d0, a_lo + 6, b_lo + 6, idesc, 1u);
              if (two) {
                umma_bf16_lohi(d1, a_lo + 1024 + 0, b_lo + 0, idesc, first);      // +128 rows x 128 B = 1024 x 16 B
                umma_bf16_lohi(d1, a_lo + 1024 + 2, b_lo + 2, idesc, 1u);
                umma_bf16_lohi(d1, a_lo + 1024 + 4, b_lo + 4, idesc, 1u);
                umma_bf16_lohi(d1, a_lo + 1024 + 6, b_lo + 6, idesc, 1u);
              }
            } else {
              for (int k = 0; k < ksteps; ++k) umma_bf16_lohi(d0, a_lo + 2 * k, b_lo + 2 * k, idesc, (first | (uint32_t)k) ? 1u : 0u);
              if (two)
                for (int k = 0; k < ksteps; ++k) umma_bf16_lohi(d1, a_lo + 1024 + 2 * k, b_lo + 2 * k, idesc, (first | (uint32_t)k) ? 1u : 0u);
            }
            umma_commit(&b_empty[sb]);
            if (j == a.taps - 1) umma_commit(&a_empty[sa]);
            if (j == a.taps - 1 && c == a.kchunks - 1) umma_commit(&acc_full[acc]);
          }
          __syncwarp();
          if (++sb == sc.nB) { sb = 0; pb ^= 1; }
        }
        if (++sa == sc.nA) { sa = 0; pa ^= 1; }
      }
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane quarter q = warp % 4, M half = (warp - 2) / 4 =====
    const int q = warp & 3, h = (warp - 2) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < sc.num_tiles; tile += gridDim.x, ++it) {
      const int bg = tile / per_bg, rem = tile - bg * per_bg;
      const int nt = rem / sc.m_tiles, mt = rem - nt * sc.m_tiles;
      const int b = bg / a.groups, g = bg - b * a.groups;
      const int n0 = nt * a.BN;
      const int t = mt * BM2 + h * 128 + q * 32 + lane;
      const int acc = sc.nacc == 2 ? (it & 1) : 0;
      const uint32_t accphase = sc.nacc == 2 ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      const bool half_ok = mt * BM2 + h * 128 < a.M;          // warp-uniform: this half holds valid rows
      const long obase = (long)b * a.o_bstride;
      const long rowflat = (long)t * a.ldo + (long)g * a.N + a.o_shift;
      const bool row_ok = t < a.M;
      float4 res[8];
      if (!ROPE && a.res != nullptr && half_ok) epi_prefetch_res(a, row_ok, obase, rowflat, n0, res);   // overlaps the main loop
      mbar_wait(&acc_full[acc], accphase);
      tc_fence_after();
      if (half_ok) {
        const uint32_t taddr = tmem_base + (uint32_t)(acc * 2 * sc.half_stride + h * sc.half_stride) + ((uint32_t)(q * 32) << 16);
        float4 res_b[8];
        auto chunk = [&](int cb, const float4 (&cur)[8], float4 (&nxt)[8]) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)cb, r);
          if (!ROPE && a.res != nullptr && cb + 32 < a.BN) epi_prefetch_res(a, row_ok, obase, rowflat, n0 + cb + 32, nxt);
          tmem_ld_wait();
          if (row_ok) epilogue_chunk32<ROPE>(a, r, cur, g, t, n0 + cb, obase, rowflat);
        };
        for (int cb = 0; cb < a.BN; cb += 64) {              // ping-pong residual buffers: no register copies
          chunk(cb, res, res_b);
          if (cb + 32 < a.BN) chunk(cb + 32, res_b, res);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

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
      if (row_ok) epilogue_chunk16(a, r, b, g, t, n0 + cb, obase, rowflat);
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

namespace {

bool use_v1() {
  static const bool v1 = [] {
    const char* e = getenv("B200TTS_GEMM");
    return e != nullptr && std::string(e) == "v1";
  }();
  return v1;
}

int sm_count() {
  static const int n = [] {
    int dev = 0, v = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    return v;
  }();
  return n;
}

TcArgs make_args(const RowGemm& p, const TcWeight& w) {
  TcArgs a;
  a.Cin = p.Cin; a.N = p.N; a.taps = p.taps; a.dil = p.dil; a.center = p.center; a.groups = p.groups; a.M = p.M;
  a.BN = w.BN; a.kchunks = ceil_div(p.Cin, BK); a.stages = 0;
  a.out = p.out; a.o_bstride = p.o_bstride; a.ldo = p.ldo; a.o_shift = p.o_shift;
  a.o_limit = p.o_limit ? p.o_limit : (long)p.M * p.ldo;
  a.out_bf16 = p.out_bf16;
  a.bias = p.bias; a.gate = p.gate; a.res = p.res; a.accumulate = p.accumulate; a.scale = p.scale; a.act = p.act;
  a.rope_cos = p.rope_cos; a.rope_sin = p.rope_sin; a.rope_cols = p.rope_cols; a.rope_rows = p.rope_rows > 0 ? p.rope_rows : 1;
  a.vt_out = p.vt_out; a.vt_col0 = p.vt_col0; a.vt_ld = p.vt_ld; a.vt_heads = p.vt_heads;
  a.out2 = p.out2;
  return a;
}

void launch_v1(const RowGemm& p, const TcWeight& w, cudaStream_t stream) {
  CUtensorMap map_a;
  tc_encode_map(&map_a, p.x, (uint64_t)p.groups * p.Cin, (uint64_t)p.Lin, (uint64_t)p.B, (uint64_t)p.ldx,
                (uint64_t)p.x_bstride, (uint32_t)BM);
  TcArgs a = make_args(p, w);
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

void launch_v2(const RowGemm& p, const TcWeight& w, cudaStream_t stream) {
  CUtensorMap map_a;
  tc_encode_map(&map_a, p.x, (uint64_t)p.groups * p.Cin, (uint64_t)p.Lin, (uint64_t)p.B, (uint64_t)p.ldx,
                (uint64_t)p.x_bstride, (uint32_t)A_BOX_ROWS);
  TcArgs a = make_args(p, w);
  Tc2Sched sc;
  sc.m_tiles = ceil_div(p.M, BM2);
  sc.n_tiles = ceil_div(p.N, w.BN);
  const long tiles = (long)p.B * p.groups * sc.m_tiles * sc.n_tiles;
  B2_CHECK(tiles < (1L << 30), "rowgemm_tc: too many tiles");
  sc.num_tiles = (int)tiles;
  const int halo = (p.taps - 1) * p.dil;
  sc.a_rows = (int)round_up(BM2 + halo, A_BOX_ROWS);
  sc.half_stride = w.BN <= 128 ? 128 : 256;
  sc.nacc = w.BN <= 128 ? 2 : 1;
  const int a_stage = sc.a_rows * 128, b_stage = w.BN * 128;
  const int bar_bytes = (2 * MAX_A_STAGES + 2 * MAX_B_STAGES + 4) * 8 + 16;
  const int budget = 227 * 1024 - 1024 - bar_bytes;
  int nA = p.taps == 1 ? MAX_A_STAGES : 2;
  if (nA > a.kchunks + 1) nA = a.kchunks + 1;
  if (nA < 1) nA = 1;
  while (nA > 1 && budget - nA * a_stage < 2 * b_stage) --nA;
  int nB = (budget - nA * a_stage) / b_stage;
  if (nB > MAX_B_STAGES) nB = MAX_B_STAGES;
  B2_CHECK(nB >= 1, "rowgemm_tc: halo tile does not fit shared memory (kernel too long / dilation too large)");
  sc.nA = nA; sc.nB = nB;
  const int smem = nA * a_stage + nB * b_stage + 1024 + bar_bytes;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    B2_CUDA(cudaFuncSetAttribute(rowgemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2_CUDA(cudaFuncSetAttribute(rowgemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  });
  const int grid = sc.num_tiles < sm_count() ? sc.num_tiles : sm_count();
  if (p.rope_cos != nullptr) {
    B2_CHECK(p.gate == nullptr && p.res == nullptr && !p.accumulate && p.act == ACT_NONE && p.out2 == nullptr,
             "rowgemm_tc: the rope epilogue takes bias only");
    rowgemm_tc2_kernel<true><<<grid, NTHREADS2, smem, stream>>>(map_a, w.map, a, sc);
  } else {
    rowgemm_tc2_kernel<false><<<grid, NTHREADS2, smem, stream>>>(map_a, w.map, a, sc);
  }
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace

void rowgemm_tc(const RowGemm& p, const TcWeight& w, cudaStream_t stream) {
  B2_CHECK(w.ready, "rowgemm_tc: tensor-core weights not prepared");
  B2_CHECK(p.Cin == w.Cin && p.N == w.N && p.taps == w.taps && p.groups == w.groups, "rowgemm_tc: weight/problem mismatch");
  B2_CHECK(p.N % 4 == 0 && p.ldo % 4 == 0 && p.o_shift % 4 == 0, "rowgemm_tc: output alignment");
  B2_CHECK(p.ldx % 8 == 0 && p.x_bstride % 8 == 0, "rowgemm_tc: A rows must be 16-byte aligned");
  B2_CHECK(p.groups == 1 || p.Cin % BK == 0, "rowgemm_tc: grouped problems need Cin % 64 == 0");
  B2_CHECK(p.M > 0 && p.B > 0, "rowgemm_tc: empty problem");
  B2_CHECK(p.rope_cos == nullptr || (p.rope_sin != nullptr && p.rope_cols % 64 == 0 && p.groups == 1 && p.B == 1),
           "rowgemm_tc: malformed rope epilogue");
  if (use_v1()) launch_v1(p, w, stream);
  else launch_v2(p, w, stream);
}

}  // namespace b200tts

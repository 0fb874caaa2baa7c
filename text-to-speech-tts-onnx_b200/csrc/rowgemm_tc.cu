// Shifted-row GEMM on the 5th-gen tensor cores (sm_100a): TMA -> 128B-swizzled shared memory ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld -> shared-memory transpose -> coalesced epilogue.
//
//   D[BM rows (time) x BN (out channels)] += A[BM x 64] . B[BN x 64]^T     per (tap, 64-channel chunk)
//
// A tiles are boxes {64 ch, 64 rows, 1 batch} of the channels-last activation tensor, fetched at row
// coordinate t0 - center*dil: rows outside [0, Lin) (negative included) and channels beyond Cin are zero-filled by
// the TMA unit, which IS the convolution's zero padding - no im2col, no halo code.
// B tiles are boxes {64 ch, BN, 1} of W[g*taps + tap][n][c].
//
// Kernel (rowgemm_tc3_kernel): persistent, one CTA per SM, BM = 256 (two M=128 accumulators that share every B tile)
// or BM = 128 (one accumulator; picked when 256-row tiles would leave SMs idle), accumulators double-buffered in TMEM
// when they fit so the epilogue of tile i overlaps the main loop of tile i+1. The A operand of a convolution is
// fetched ONCE per 64-channel chunk as a (BM + (taps-1)*dil)-row halo tile; every tap is an MMA whose A descriptor
// starts (tap*dil) rows further down that tile (plain descriptor start-address shift; the swizzle XOR is a function of
// the absolute smem address).
// 320 threads: warps 0-7 = epilogue, warp 8 = TMA producer, warp 9 = TMEM owner + MMA issuer.
//
// Epilogue (r01c): the first two versions had every thread own one output ROW and walk its columns, which (a) made each
// warp-wide global access touch 32 different 128-byte lines, and (b) unrolled into 11.6k SASS instructions executed
// once per tile, so the warps sat in instruction-fetch (stall_no_inst 25 %) and L2-latency (stall_long_sb 39 %) stalls:
// 70-120k cycles per tile against an 8k-cycle main loop (ncu, profiles/r01). Now each warp moves a 32x32 fp32 block
// TMEM -> registers -> swizzled shared memory, re-reads it with 8 lanes across a row's 32 columns (4 rows per
// instruction) and applies bias / activation / gate / residual / RoPE on that layout, so every global load and store
// is a run of full 128-byte (fp32) or 64-byte (bf16) row segments; the residual of the next block is prefetched while
// the current one is processed; the loop over blocks is not unrolled (~300 instructions per block).
#include "rowgemm_tc.cuh"

#include <cuda_fp16.h>

#include <cstdlib>
#include <mutex>
#include <string>

#include "rowgemm_tc_dev.cuh"

namespace b200tts {

namespace {

// A tap's A tile starts (tap*dil) rows = (tap*dil)*128 bytes into the halo tile, i.e. generally NOT on a 1024-byte
// swizzle-atom boundary. Measured on B200 (tools/debug_conv.py): the tensor core applies the 128B-swizzle XOR to the
// absolute shared-memory address bits, exactly as the TMA unit did when it wrote the tile, so the plain descriptor with
// the shifted start address is correct and the matrix-base-offset field must stay 0 (setting it to the row phase
// (addr >> 7) & 7 double-applies the rotation and corrupts every tap whose shift is not a multiple of 8 rows).

// BRES: the whole weight tensor stays in shared memory for the life of the CTA (thin convolutions).
template <int KIND, int ACT, bool BRES>
__global__ void __launch_bounds__(NTHREADS3, 1) rowgemm_tc3_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_b,
                                                                   const TcArgs a, const Tc3Sched sc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_stage_bytes = sc.a_rows * 128;
  const int b_stage_bytes = a.BN * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + sc.nA * a_stage_bytes;
  float* smem_epi = reinterpret_cast<float*>(smem_b + sc.nB * b_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_epi) + EPI_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + MAX_A_STAGES;
  uint64_t* b_full = a_empty + MAX_A_STAGES;
  uint64_t* b_empty = b_full + MAX_B_STAGES;
  uint64_t* acc_full = b_empty + MAX_B_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == WARP_TMA && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < sc.nA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < (BRES ? 1 : sc.nB); ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();                                          // after the TMEM allocation (common.cuh)
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_bg = sc.m_tiles * sc.n_tiles;
  const int halo_lo = a.center * a.dil;
  const int halves = sc.halves;
  const uint32_t acc_stride = (uint32_t)(halves * sc.half_stride);

  if (warp == WARP_TMA) {
    if (lane == 0) {
      // ===== A producer (activations). A plain GEMM needs an A and a B stage every 512 tensor cycles, and ONE thread issuing
      // both (2 barrier waits, 2 expect_tx, 3 TMA ops per chunk) was the bottleneck of the batched DiT GEMMs: the MMA warp
      // spun on full barriers with the tensor pipe 50 % active (ncu r01q). A and B now have a producer thread each. =====
      int sa = 0; uint32_t pa = 0;
      const int a_boxes = sc.a_rows / sc.a_box;
      for (int tile = blockIdx.x; tile < sc.num_tiles; tile += gridDim.x) {
        const int bg = tile / per_bg, rem = tile - bg * per_bg;
        const int nt = rem / sc.m_tiles, mt = rem - nt * sc.m_tiles;
        const int b = bg / a.groups, g = bg - b * a.groups;
        const int t0 = mt * sc.bm, n0 = nt * a.BN;
        if (KIND == EPI_STD && a.o_shift == 0 && (a.res != nullptr || (a.accumulate && !a.out_bf16))) {
          // The epilogue reads this tile's residual (and, when accumulating, the old output) one 32x32 block ahead of its
          // use, which covers an L2 hit but not a DRAM miss (conv2 of the thin BigVGAN stages ran at 2 TB/s, ncu r01l).
          // The producer runs a tile or two ahead of the epilogue: it pulls those tiles into L2 here, for free.
          int rows = a.M - t0; if (rows > sc.bm) rows = sc.bm;
          int cols = a.N - n0; if (cols > a.BN) cols = a.BN;
          const long toff = (long)b * a.o_bstride + (long)t0 * a.ldo + (long)g * a.N + n0;
          for (int which = 0; which < 2; ++which) {
            const float* rp = which == 0 ? a.res : (a.accumulate && !a.out_bf16 ? reinterpret_cast<const float*>(a.out) : nullptr);
            if (rp == nullptr) continue;
            rp += toff;
            if (cols == a.ldo) {                                // one N tile spans the row: the tile is contiguous
              const long bytes = (long)rows * a.ldo * 4;
              for (long off = 0; off < bytes; off += 65536)
                prefetch_l2_bulk(reinterpret_cast<const char*>(rp) + off, (uint32_t)(bytes - off < 65536 ? bytes - off : 65536));
            } else if (sc.num_tiles > (int)gridDim.x) {          // (a one-tile CTA has no lead time to gain)
              for (int r = 0; r < rows; ++r) prefetch_l2_bulk(rp + (long)r * a.ldo, (uint32_t)cols * 4u);
            }
          }
        }
        for (int c = 0; c < a.kchunks; ++c) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], (uint32_t)a_stage_bytes);
          for (int rbx = 0; rbx < a_boxes; ++rbx)
            tma_load_3d(smem_a + sa * a_stage_bytes + rbx * (sc.a_box * 128), &map_a, &a_full[sa], g * a.Cin + c * BK,
                        t0 - halo_lo + rbx * sc.a_box, b);
          if (++sa == sc.nA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == WARP_TMA_B) {
    if (lane == 0) {
      // ===== B producer (weights) =====
      if (BRES) {                       // small convolutions: every (chunk, tap) weight tile is fetched once per CTA
        mbar_expect_tx(&b_full[0], (uint32_t)(a.kchunks * a.taps * b_stage_bytes));
        for (int c = 0; c < a.kchunks; ++c)
          for (int j = 0; j < a.taps; ++j)
            tma_load_3d(smem_b + (c * a.taps + j) * b_stage_bytes, &map_b, &b_full[0], c * BK, 0, j);
      } else {
        int sb = 0; uint32_t pb = 0;
        for (int tile = blockIdx.x; tile < sc.num_tiles; tile += gridDim.x) {
          const int bg = tile / per_bg, rem = tile - bg * per_bg;
          const int nt = rem / sc.m_tiles;
          const int g = bg % a.groups;
          const int n0 = nt * a.BN;
          for (int c = 0; c < a.kchunks; ++c)
            for (int j = 0; j < a.taps; ++j) {
              mbar_wait(&b_empty[sb], pb ^ 1);
              mbar_expect_tx(&b_full[sb], (uint32_t)b_stage_bytes);
              tma_load_3d(smem_b + sb * b_stage_bytes, &map_b, &b_full[sb], c * BK, n0, g * a.taps + j);
              if (++sb == sc.nB) { sb = 0; pb ^= 1; }
            }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===== MMA issuer: the whole warp walks the (warp-uniform) loops, one elected lane issues =====
    const uint32_t idesc = idesc_f16kind(128, a.BN, a.f16);
    const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a)), b_lo0 = desc_lo_sw128(smem_u32(smem_b));
    const uint32_t a_stage_lo = (uint32_t)a_stage_bytes >> 4, b_stage_lo = (uint32_t)b_stage_bytes >> 4;
    const uint32_t tap_lo = (uint32_t)a.dil * 8u;                   // dil rows x 128 B, in 16-byte units
    const uint32_t hs = (uint32_t)sc.half_stride;
    const int kchunks = a.kchunks, taps = a.taps, Cin = a.Cin, nA = sc.nA, nB = sc.nB;
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < sc.num_tiles; tile += gridDim.x, ++it) {
      const int rem = tile % per_bg;
      const int mt = rem % sc.m_tiles;
      int nh = (a.M - mt * sc.bm + 127) >> 7;                        // valid 128-row halves of a ragged last tile
      if (nh > halves) nh = halves;
      const int acc = sc.nacc == 2 ? (it & 1) : 0;
      const uint32_t accphase = sc.nacc == 2 ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      mbar_wait(&acc_empty[acc], accphase ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)acc * acc_stride;
      for (int c = 0; c < kchunks; ++c) {
        int ksteps = (Cin - c * BK + UMMA_K - 1) / UMMA_K;
        if (ksteps > BK / UMMA_K) ksteps = BK / UMMA_K;
        const int mode = ksteps == 4 ? nh : 0;                          // full chunk, 1 / 2 / 4 halves: the hot, fully unrolled path
        mbar_wait(&a_full[sa], pa);
        uint32_t a_lo = a_lo0 + (uint32_t)sa * a_stage_lo;
        for (int j = 0; j < taps; ++j, a_lo += tap_lo) {
          if (BRES) { if (it == 0 && c == 0 && j == 0) mbar_wait(&b_full[0], 0); }
          else mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + (uint32_t)(BRES ? c * taps + j : sb) * b_stage_lo;
          const uint32_t accum = (c > 0 || j > 0) ? 1u : 0u;
          if (elect_one()) {
            // halves innermost: consecutive MMAs go to different accumulators, so the chain of dependent accumulations
            // into one accumulator never stalls the tensor pipe (matters when an MMA is short: small BN). The common
            // shapes are fully unrolled and every loop bound lives in a register: ONE thread issues every MMA of the
            // CTA, and a generic double loop here cost 40 % of the GEMM throughput (sweep_e.log).
            bool done = true;
            if (mode == 1) issue_tap<1, 4>(d0, hs, a_lo, b_lo, idesc, accum);
            else if (mode == 2) issue_tap<2, 4>(d0, hs, a_lo, b_lo, idesc, accum);
            else if (mode == 4) issue_tap<4, 4>(d0, hs, a_lo, b_lo, idesc, accum);
            else if (nh == 1) done = issue_tap_ks<1>(ksteps, d0, hs, a_lo, b_lo, idesc, accum);      // ragged last chunk
            else if (nh == 2) done = issue_tap_ks<2>(ksteps, d0, hs, a_lo, b_lo, idesc, accum);
            else if (nh == 4) done = issue_tap_ks<4>(ksteps, d0, hs, a_lo, b_lo, idesc, accum);
            else done = false;
            if (!done)
              for (int k = 0; k < ksteps; ++k)
                for (int h = 0; h < nh; ++h)
                  umma_bf16_lohi(d0 + (uint32_t)h * hs, a_lo + (uint32_t)h * 1024u + 2u * k, b_lo + 2u * k, idesc,
                                 (accum | (uint32_t)k) ? 1u : 0u);      // +128 rows x 128 B = 1024 x 16 B per half
            if (!BRES) umma_commit(&b_empty[sb]);
            if (j == taps - 1) umma_commit(&a_empty[sa]);
            if (j == taps - 1 && c == kchunks - 1) umma_commit(&acc_full[acc]);
          }
          __syncwarp();
          if (!BRES && ++sb == nB) { sb = 0; pb ^= 1; }
        }
        if (++sa == nA) { sa = 0; pa ^= 1; }
      }
    }
  } else {
    // ===== epilogue: warps 0..7; TMEM lane quarter q = warp % 4; e = warp / 4 selects the odd / even 32-column
    // blocks (1 half), the M half (2 halves) or the pair of halves {e, e+2} (4 halves) =====
    const int q = warp & 3, e = warp >> 2;
    float* stg = smem_epi + warp * (EPI_STAGE_BYTES / 4);
    const int cb_first = halves == 1 ? e * 32 : 0, cb_step = halves == 1 ? 64 : 32;
    const int h_first = halves == 1 ? 0 : e;
    const bool has_res = KIND == EPI_STD && a.res != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < sc.num_tiles; tile += gridDim.x, ++it) {
      const int bg = tile / per_bg, rem = tile - bg * per_bg;
      const int nt = rem / sc.m_tiles, mt = rem - nt * sc.m_tiles;
      const int b = bg / a.groups, g = bg - b * a.groups;
      EpiPos p;
      p.sub = lane >> 3; p.c4 = lane & 7;
      p.n0 = nt * a.BN;
      p.t_row0 = mt * sc.bm + h_first * 128 + q * 32;
      p.obase = (long)b * a.o_bstride;
      p.gshift = (long)g * a.N + a.o_shift;
      float4 res[8];
      if (has_res && p.t_row0 < a.M) epi_load_res(a, p, cb_first, res);      // overlaps the main loop
      const int acc = sc.nacc == 2 ? (it & 1) : 0;
      const uint32_t accphase = sc.nacc == 2 ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      mbar_wait(&acc_full[acc], accphase);
      tc_fence_after();
      {
        const uint32_t taddr = tmem_base + (uint32_t)acc * acc_stride + (uint32_t)(h_first * sc.half_stride) + ((uint32_t)(q * 32) << 16);
        epilogue_warp<KIND, ACT>(a, p, taddr, 2u * (uint32_t)sc.half_stride, halves == 4 ? 2 : 1, stg, lane, g, cb_first, cb_step, res);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =============================================================================================
// CTA-pair variant (cta_group::2): the two SMs of a TPC run ONE 256 x BN tile. Each CTA fetches its 128 rows of A (plus the
// tap halo) and HALF of every B tile (BN/2 weight rows); the leader's single MMA thread issues M=256 tcgen05.mma
// instructions that read both CTAs' shared memory and write 128 accumulator rows into each CTA's TMEM.
// Why: the one-CTA kernel is bound by the L2 -> SM fabric (~6300 B/clk chip-wide, ~42.6 B/clk per SM, B300_MICROARCH.md):
// a 128x256 tile needs 64 B/clk per SM for the weights of a convolution (-> 67 % of the tensor peak, exactly the 1.53
// PFLOP/s measured on the C=768 convs) and 94 B/clk for a plain GEMM (-> 45 %, the ~1.0 PFLOP/s of the batched DiT GEMMs).
// Halving the B bytes each SM receives puts both under the cap.
// Barriers: every "full" barrier lives in the leader (both CTAs' TMA loads signal it through the peer-bit-masked
// address), every "empty" barrier is per CTA and is released by a multicast tcgen05.commit; the accumulator-empty
// barrier of the leader collects the epilogue warps of both CTAs.
// =============================================================================================
template <int KIND, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS3, 1)
rowgemm_tc2sm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcArgs a,
                     const Tc3Sched sc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_stage_bytes = sc.a_rows * 128;
  const int b_stage_bytes = (a.BN / 2) * 128;            // this CTA's half of a B tile
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + sc.nA * a_stage_bytes;
  float* smem_epi = reinterpret_cast<float*>(smem_b + sc.nB * b_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_epi) + EPI_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + MAX_A_STAGES;
  uint64_t* b_full = a_empty + MAX_A_STAGES;
  uint64_t* b_empty = b_full + MAX_B_STAGES;
  uint64_t* acc_full = b_empty + MAX_B_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == WARP_TMA && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < sc.nA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < sc.nB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 16); }
    fence_barrier_init();
  }
  cluster_sync_all();                                     // the peer's barriers exist before anything signals them
  if (warp == WARP_MMA) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  cluster_sync_all();                                     // both allocations are done before the leader's MMAs write the peer's TMEM
  tc_fence_after();
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_bg = sc.m_tiles * sc.n_tiles;             // pair tiles (256 rows x BN) per (batch, group)
  const int halo_lo = a.center * a.dil;
  const uint32_t acc_stride = (uint32_t)sc.half_stride;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      // ===== A producer (both CTAs): this CTA's 128 rows (+ tap halo) =====
      int sa = 0; uint32_t pa = 0;
      const int a_boxes = sc.a_rows / sc.a_box;
      for (int tile = cluster_id; tile < sc.num_tiles; tile += nclusters) {
        const int bg = tile / per_bg, rem = tile - bg * per_bg;
        const int nt = rem / sc.m_tiles, mt = rem - nt * sc.m_tiles;
        const int b = bg / a.groups, g = bg - b * a.groups;
        const int t0 = mt * 256 + (int)rank * 128;
        if (KIND == EPI_STD && a.o_shift == 0 && a.res != nullptr && t0 < a.M) {      // residual rows of this CTA -> L2
          int rows = a.M - t0; if (rows > 128) rows = 128;
          int cols = a.N - nt * a.BN; if (cols > a.BN) cols = a.BN;
          const float* rp = a.res + (long)b * a.o_bstride + (long)t0 * a.ldo + (long)g * a.N + nt * a.BN;
          if (cols == a.ldo) {
            const long bytes = (long)rows * a.ldo * 4;
            for (long off = 0; off < bytes; off += 65536)
              prefetch_l2_bulk(reinterpret_cast<const char*>(rp) + off, (uint32_t)(bytes - off < 65536 ? bytes - off : 65536));
          } else {
            for (int r = 0; r < rows; ++r) prefetch_l2_bulk(rp + (long)r * a.ldo, (uint32_t)cols * 4u);
          }
        }
        for (int c = 0; c < a.kchunks; ++c) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (leader) mbar_expect_tx(&a_full[sa], 2u * (uint32_t)a_stage_bytes);
          for (int rbx = 0; rbx < a_boxes; ++rbx)
            tma_load_3d_2sm(smem_a + sa * a_stage_bytes + rbx * (sc.a_box * 128), &map_a, &a_full[sa], g * a.Cin + c * BK,
                            t0 - halo_lo + rbx * sc.a_box, b);
          if (++sa == sc.nA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == WARP_TMA_B) {
    if (lane == 0) {
      // ===== B producer (both CTAs): this CTA's half of every weight tile =====
      int sb = 0; uint32_t pb = 0;
      for (int tile = cluster_id; tile < sc.num_tiles; tile += nclusters) {
        const int bg = tile / per_bg, rem = tile - bg * per_bg;
        const int nt = rem / sc.m_tiles;
        const int g = bg % a.groups;
        const int n0 = nt * a.BN + (int)rank * (a.BN / 2);
        for (int c = 0; c < a.kchunks; ++c)
          for (int j = 0; j < a.taps; ++j) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            if (leader) mbar_expect_tx(&b_full[sb], 2u * (uint32_t)b_stage_bytes);
            tma_load_3d_2sm(smem_b + sb * b_stage_bytes, &map_b, &b_full[sb], c * BK, n0, g * a.taps + j);
            if (++sb == sc.nB) { sb = 0; pb ^= 1; }
          }
      }
    }
  } else if (warp == WARP_MMA) {
    if (leader) {
      // ===== MMA issuer (leader CTA only): M = 256 across the pair =====
      const uint32_t idesc = idesc_f16kind(256, a.BN, a.f16);
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a)), b_lo0 = desc_lo_sw128(smem_u32(smem_b));
      const uint32_t a_stage_lo = (uint32_t)a_stage_bytes >> 4, b_stage_lo = (uint32_t)b_stage_bytes >> 4;
      const uint32_t tap_lo = (uint32_t)a.dil * 8u;
      const int kchunks = a.kchunks, taps = a.taps, Cin = a.Cin, nA = sc.nA, nB = sc.nB;
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = cluster_id; tile < sc.num_tiles; tile += nclusters, ++it) {
        const int acc = it & 1;
        const uint32_t accphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&acc_empty[acc], accphase ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)acc * acc_stride;
        for (int c = 0; c < kchunks; ++c) {
          int ksteps = (Cin - c * BK + UMMA_K - 1) / UMMA_K;
          if (ksteps > BK / UMMA_K) ksteps = BK / UMMA_K;
          mbar_wait(&a_full[sa], pa);
          uint32_t a_lo = a_lo0 + (uint32_t)sa * a_stage_lo;
          for (int j = 0; j < taps; ++j, a_lo += tap_lo) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_stage_lo;
            const uint32_t accum = (c > 0 || j > 0) ? 1u : 0u;
            if (elect_one()) {
              if (ksteps == 4) {
                umma2_bf16_lohi(d0, a_lo + 0, b_lo + 0, idesc, accum);
                umma2_bf16_lohi(d0, a_lo + 2, b_lo + 2, idesc, 1u);
                umma2_bf16_lohi(d0, a_lo + 4, b_lo + 4, idesc, 1u);
                umma2_bf16_lohi(d0, a_lo + 6, b_lo + 6, idesc, 1u);
              } else {
                for (int k = 0; k < ksteps; ++k) umma2_bf16_lohi(d0, a_lo + 2u * k, b_lo + 2u * k, idesc, (accum | (uint32_t)k) ? 1u : 0u);
              }
              umma2_commit_mc(&b_empty[sb]);
              if (j == taps - 1) umma2_commit_mc(&a_empty[sa]);
              if (j == taps - 1 && c == kchunks - 1) umma2_commit_mc(&acc_full[acc]);
            }
            __syncwarp();
            if (++sb == nB) { sb = 0; pb ^= 1; }
          }
          if (++sa == nA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the pair tile; e = warp / 4 takes the even / odd 32-column blocks =====
    const int q = warp & 3, e = warp >> 2;
    float* stg = smem_epi + warp * (EPI_STAGE_BYTES / 4);
    const int cb_first = e * 32, cb_step = 64;
    const bool has_res = KIND == EPI_STD && a.res != nullptr;
    int it = 0;
    for (int tile = cluster_id; tile < sc.num_tiles; tile += nclusters, ++it) {
      const int bg = tile / per_bg, rem = tile - bg * per_bg;
      const int nt = rem / sc.m_tiles, mt = rem - nt * sc.m_tiles;
      const int b = bg / a.groups, g = bg - b * a.groups;
      EpiPos p;
      p.sub = lane >> 3; p.c4 = lane & 7;
      p.n0 = nt * a.BN;
      p.t_row0 = mt * 256 + (int)rank * 128 + q * 32;
      p.obase = (long)b * a.o_bstride;
      p.gshift = (long)g * a.N + a.o_shift;
      float4 res[8];
      if (has_res && p.t_row0 < a.M) epi_load_res(a, p, cb_first, res);
      const int acc = it & 1;
      const uint32_t accphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_full[acc], accphase);
      tc_fence_after();
      {
        const uint32_t taddr = tmem_base + (uint32_t)acc * acc_stride + ((uint32_t)(q * 32) << 16);
        epilogue_warp<KIND, ACT>(a, p, taddr, 0u, 1, stg, lane, g, cb_first, cb_step, res);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
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

// 8-bit operands (e4m3): dims {d0 bytes (contiguous), d1, d2}, byte strides, box {128, box1, 1}, SWIZZLE_128B
void tc_encode_map_u8(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2, uint32_t box1) {
  B2_CHECK(((uintptr_t)base & 15) == 0 && ld1 % 16 == 0 && ld2 % 16 == 0, "TMA base / strides must be multiples of 16 bytes");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld1, ld2};
  cuuint32_t box[3] = {128u, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled (u8) failed with CUresult " + std::to_string((int)r));
}

void tc_encode_map2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t ld1, uint32_t box0,
                     uint32_t box1) {
  B2_CHECK(elem_bytes == 2 || elem_bytes == 4, "tc_encode_map2d: element size");
  B2_CHECK(((uintptr_t)base & 15) == 0 && (ld1 * elem_bytes) % 16 == 0, "TMA base / stride must be 16-byte aligned");
  const uint32_t inner = box0 * (uint32_t)elem_bytes;
  B2_CHECK(inner == 128 || inner == 64, "tc_encode_map2d: box rows must be 128 or 64 bytes");
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {ld1 * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                            const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled (2-D) failed with CUresult " + std::to_string((int)r));
}

namespace {

__global__ void cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long n, int f16) {
  pdl_trigger();
  pdl_wait();
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + i));
    *reinterpret_cast<uint2*>(out + i) = pack16x4(v.x, v.y, v.z, v.w, f16);
  } else {
    for (long k = i; k < n; ++k) reinterpret_cast<uint16_t*>(out)[k] = pack16(in[k], f16);
  }
}

__global__ void uncast_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long n, int f16) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = f16 ? __half2float(reinterpret_cast<const __half*>(in)[i]) : __bfloat162float(in[i]);
}

__global__ void cast_pad_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long rows, int C, int ldo, int f16) {
  pdl_trigger();
  pdl_wait();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ldo) return;
  const long r = i / ldo;
  const int c = (int)(i - r * ldo);
  reinterpret_cast<uint16_t*>(out)[i] = pack16(c < C ? in[r * C + c] : 0.f, f16);
}

// the same for the 2 * Ntot DiT rows of a ragged batch: row r = (sequence, position) reads the token row of its utterance, which
// both CFG sequences of the utterance share (sequence 2u starts at DiT row 2 * tok_off[u])
__global__ void cast_pad_rows_ragged_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, const int2* __restrict__ rowinfo,
                                            const int* __restrict__ seq_off, long rows, int C, int ldo, int f16) {
  pdl_trigger();
  pdl_wait();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ldo) return;
  const long r = i / ldo;
  const int c = (int)(i - r * ldo);
  const int2 ri = rowinfo[r];
  const long tok = (long)(seq_off[ri.x & ~1] >> 1) + ri.y;
  reinterpret_cast<uint16_t*>(out)[i] = pack16(c < C ? in[tok * C + c] : 0.f, f16);
}

}  // namespace

void cast_pad_rows_ragged(const float* in, __nv_bfloat16* out, const int2* rowinfo, const int* seq_off, long rows, int C, int ldo,
                          cudaStream_t s, int f16) {
  if (rows <= 0) return;
  launch_pdl(cast_pad_rows_ragged_kernel, dim3(ceil_div(rows * ldo, 256)), dim3(256), 0, s, in, out, rowinfo, seq_off, rows, C, ldo, f16);
  B2_LAUNCH_CHECK(); count_launch();
}

void cast_f32_to_bf16(const float* in, __nv_bfloat16* out, long n, cudaStream_t s, int f16) {
  if (n <= 0) return;
  B2_CHECK(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0, "cast alignment");
  launch_pdl(cast_kernel, dim3(ceil_div(ceil_div(n, 4), 256)), dim3(256), 0, s, in, out, n, f16);
  B2_LAUNCH_CHECK(); count_launch();
}

void cast_bf16_to_f32(const __nv_bfloat16* in, float* out, long n, cudaStream_t s, int f16) {
  if (n <= 0) return;
  uncast_kernel<<<ceil_div(n, 256), 256, 0, s>>>(in, out, n, f16);
  B2_LAUNCH_CHECK(); count_launch();
}

void cast_pad_f32_to_bf16(const float* in, __nv_bfloat16* out, long rows, int C, int ldo, cudaStream_t s, int f16) {
  if (rows <= 0) return;
  launch_pdl(cast_pad_kernel, dim3(ceil_div(rows * ldo, 256)), dim3(256), 0, s, in, out, rows, C, ldo, f16);
  B2_LAUNCH_CHECK(); count_launch();
}

void tc_weight_from_f32(TcWeight& tw, const float* w_gjnc, int groups, int taps, int N, int Cin, cudaStream_t s, int f16) {
  tw.Cin = Cin; tw.N = N; tw.taps = taps; tw.groups = groups; tw.f16 = f16;
  tw.ldc = (int)round_up(Cin, 8);
  const long rows = (long)groups * taps * N;
  tw.w.alloc((size_t)rows * tw.ldc);
  cast_pad_f32_to_bf16(w_gjnc, tw.w.p, rows, Cin, tw.ldc, s, f16);
  tw.ready = true;
}

namespace {

int sm_count() {
  static const int n = [] {
    int dev = 0, v = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    return v;
  }();
  return n;
}

// Tile shape (bm, BN) for a problem of `units` = B*groups independent (M x N) outputs on `sms` SMs.
// Cost model fitted to the tile sweep of profiles/r01/sweep_gemm_c.log (B200, 148 SMs):
//   time ~ waves x (bm x BN) x f,   f = max(1, (1/bm + 1/BN) / (1/128 + 1/256)) x mma(BN)
// The first factor is the L2->SMEM operand traffic per FLOP relative to a 128x256 tile (smaller tiles are L2-bandwidth
// bound), the second the shared-memory read rate of the MMA itself: an M=128 UMMA re-reads its B operand, so N=128 needs
// 128 B/clk of shared memory against 96 B/clk at N=256 (measured 1.2x slower per FLOP, N=192 1.07x).
// A single accumulator stage (no epilogue / main-loop overlap) is charged when a CTA runs more than one tile.
struct TileShape { int halves, bn; };
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
// Experiment knobs (tools/sweep_gemm.sh, tools/test_2sm.sh), read once per process.
struct TcEnv { int bm, bn, na, bres, two_sm; };
const TcEnv& tc_env() {
  static const TcEnv e{env_int("B200TTS_BM", 0), env_int("B200TTS_BN", 0), env_int("B200TTS_NA", 0), env_int("B200TTS_BRES", 1),
                       env_int("B200TTS_2SM", -1)};
  return e;
}
int half_stride_of(int bn) { return bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }

TileShape pick_tile(int M, int N, long units, int sms, bool rope) {
  int cands[8]; int nc = 0;
  if (N < 64) {
    cands[nc++] = (int)round_up(N, 16);                    // one N tile; rows beyond N are TMA zero fill
  } else {
    for (int bn : {256, 192, 128, 96, 64})
      if (N % bn == 0 || (bn == 128 && N % 64 != 0 && N % 96 != 0)) cands[nc++] = bn;     // 128: ragged-tail fallback
  }
  TileShape best{1, cands[0]};
  double best_cost = 1e300;
  for (int ci = 0; ci < nc; ++ci) {
    const int bn = cands[ci];
    if (rope && bn % 64 != 0) continue;
    for (int halves : {1, 2, 4}) {
      const int hs = half_stride_of(bn);
      if (halves * hs > 512) continue;
      if (halves == 4 && bn > 64) continue;                // 512-row tiles only for the thin (HBM-bound) convolutions
      const int bm = 128 * halves;
      const long tiles = units * ceil_div(M, bm) * ceil_div(N, bn);
      const long waves = (tiles + sms - 1) / sms;
      const double traffic = (1.0 / bm + 1.0 / bn) / (1.0 / 128 + 1.0 / 256);
      const double mma = bn >= 256 ? 1.0 : bn >= 192 ? 1.07 : 1.2;
      const bool single_acc = 2 * halves * hs > 512;
      double cost = (double)waves * bm * bn * (traffic > 1.0 ? traffic : 1.0) * mma;
      if (single_acc && waves > 1) cost *= 1.1;
      if (cost < best_cost) { best_cost = cost; best = TileShape{halves, bn}; }
    }
  }
  // experiment overrides (tools/sweep_gemm.sh)
  const int ebm = tc_env().bm, ebn = tc_env().bn;
  if (ebm && ebn && (ebm == 128 || ebm == 256 || ebm == 512) && ebn >= 16 && ebn <= 256 && ebn % 16 == 0 && (!rope || ebn % 64 == 0) &&
      (ebm / 128) * half_stride_of(ebn) <= 512)
    best = TileShape{ebm / 128, ebn};
  return best;
}

TcArgs make_args(const RowGemm& p, int BN) {
  TcArgs a;
  a.Cin = p.Cin; a.N = p.N; a.taps = p.taps; a.dil = p.dil; a.center = p.center; a.groups = p.groups; a.M = p.M;
  a.BN = BN; a.kchunks = ceil_div(p.Cin, BK);
  a.out = p.out; a.o_bstride = p.o_bstride; a.ldo = p.ldo; a.o_shift = p.o_shift;
  a.o_limit = p.o_limit ? p.o_limit : (long)p.M * p.ldo;
  a.out_bf16 = p.out_bf16;
  a.bias = p.bias; a.gate = p.gate; a.res = p.res; a.accumulate = p.accumulate; a.scale = p.scale;
  a.rope_cs = p.rope_cs; a.rope_cols = p.rope_cols; a.rope_rows = p.rope_rows > 0 ? p.rope_rows : 1;
  a.vt_out = p.vt_out; a.vt_col0 = p.vt_col0; a.vt_ld = p.vt_ld; a.vt_heads = p.vt_heads;
  a.out2 = p.out2;
  a.f16 = p.f16;
  a.rowinfo = p.rowinfo;
  return a;
}

template <int KIND, int ACT, bool BRES>
void launch_one(int grid, int smem, cudaStream_t stream, const CUtensorMap& map_a, const CUtensorMap& map_b, const TcArgs& a,
                const Tc3Sched& sc) {
  static PerDeviceOnce once;
  if (once.first())
    B2_CUDA(cudaFuncSetAttribute(rowgemm_tc3_kernel<KIND, ACT, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  launch_pdl(rowgemm_tc3_kernel<KIND, ACT, BRES>, dim3(grid), dim3(NTHREADS3), (size_t)smem, stream, map_a, map_b, a, sc);
}
template <int KIND, int ACT>
void launch_kernel(int grid, int smem, cudaStream_t stream, const CUtensorMap& map_a, const CUtensorMap& map_b, const TcArgs& a,
                   const Tc3Sched& sc) {
  if (sc.bres) launch_one<KIND, ACT, true>(grid, smem, stream, map_a, map_b, a, sc);
  else launch_one<KIND, ACT, false>(grid, smem, stream, map_a, map_b, a, sc);
}

template <int KIND, int ACT>
void launch_2sm(int grid, int smem, cudaStream_t stream, const CUtensorMap& map_a, const CUtensorMap& map_b, const TcArgs& a,
                const Tc3Sched& sc) {
  static PerDeviceOnce once;
  if (once.first())
    B2_CUDA(cudaFuncSetAttribute(rowgemm_tc2sm_kernel<KIND, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  launch_pdl(rowgemm_tc2sm_kernel<KIND, ACT>, dim3(grid), dim3(NTHREADS3), (size_t)smem, stream, map_a, map_b, a, sc);
}

// CTA-pair launch: 256 x BN pair tiles, BN in {256, 192, 128} dividing N. Returns false when the problem is not eligible.
bool try_launch_2sm(const RowGemm& p, const TcWeight& w, cudaStream_t stream, bool rope) {
  const int mode = tc_env().two_sm;                       // B200TTS_2SM: -1 auto, 0 off, 1 force when eligible
  if (mode == 0) return false;
  int bn = 0;
  for (int c : {256, 192, 128}) if (p.N % c == 0) { bn = c; break; }
  if (bn == 0 || (rope && bn % 64 != 0)) return false;
  const int sms = sm_count();
  if (sms < 2) return false;
  const int nclusters = sms / 2;
  const int m_tiles = ceil_div(p.M, 256), n_tiles = p.N / bn;
  const long tiles = (long)p.B * p.groups * m_tiles * n_tiles;
  const int kchunks = ceil_div(p.Cin, BK);
  // auto (test_2sm_b.log, same box): the pair kernel wins where the main loop dominates -- convolutions and GEMMs without a
  // residual operand (+5..17 %) -- and loses a few % where the fp32-residual epilogue is the critical stage; it needs at least
  // two full waves of pair tiles and a main loop long enough to amortise the cluster prologue
  if (mode < 0 && !(p.res == nullptr && tiles >= 2L * nclusters && (long)kchunks * p.taps >= 8)) return false;
  CUtensorMap map_a, map_b;
  tc_encode_map(&map_a, p.x, (uint64_t)p.groups * p.Cin, (uint64_t)p.Lin, (uint64_t)p.B, (uint64_t)p.ldx,
                (uint64_t)p.x_bstride, (uint32_t)A_BOX_ROWS);
  tc_encode_map(&map_b, w.w.p, (uint64_t)w.Cin, (uint64_t)w.N, (uint64_t)w.groups * w.taps, (uint64_t)w.ldc,
                (uint64_t)w.N * w.ldc, (uint32_t)(bn / 2));
  TcArgs a = make_args(p, bn);
  Tc3Sched sc;
  sc.bm = 256; sc.halves = 1;
  sc.m_tiles = m_tiles; sc.n_tiles = n_tiles;
  B2_CHECK(tiles < (1L << 30), "rowgemm_tc: too many tiles");
  sc.num_tiles = (int)tiles;
  const int halo = (p.taps - 1) * p.dil;
  sc.a_rows = (int)round_up(128 + halo, A_BOX_ROWS);
  sc.a_box = sc.a_rows % 256 == 0 ? 256 : sc.a_rows % 128 == 0 ? 128 : 64;
  tc_encode_map(&map_a, p.x, (uint64_t)p.groups * p.Cin, (uint64_t)p.Lin, (uint64_t)p.B, (uint64_t)p.ldx,
                (uint64_t)p.x_bstride, (uint32_t)sc.a_box);
  sc.half_stride = 256;                                   // two accumulator stages of up to 256 columns
  sc.nacc = 2; sc.bres = 0;
  const int a_stage = sc.a_rows * 128, b_stage = (bn / 2) * 128;
  const int bar_bytes = (2 * MAX_A_STAGES + 2 * MAX_B_STAGES + 4) * 8 + 16;
  const int budget = 227 * 1024 - 1024 - bar_bytes - EPI_BYTES;
  int nA, nB;
  if (p.taps == 1) {
    nA = budget / (a_stage + b_stage);
    if (nA > MAX_A_STAGES) nA = MAX_A_STAGES;
    nB = nA;
  } else {
    nA = 2;
    if (budget - nA * a_stage < 2 * b_stage) return false;
    nB = (budget - nA * a_stage) / b_stage;
    if (nB > MAX_B_STAGES) nB = MAX_B_STAGES;
  }
  if (nA < 1 || nB < 1) return false;
  sc.nA = nA; sc.nB = nB;
  const int smem = nA * a_stage + nB * b_stage + EPI_BYTES + 1024 + bar_bytes;
  long want = 2 * tiles;
  const int grid = (int)(want < 2L * nclusters ? want : 2L * nclusters);
  if (rope) {
    launch_2sm<EPI_ROPE, ACT_NONE>(grid, smem, stream, map_a, map_b, a, sc);
  } else {
    switch (p.act) {
      case ACT_NONE: launch_2sm<EPI_STD, ACT_NONE>(grid, smem, stream, map_a, map_b, a, sc); break;
      case ACT_GELU_TANH: launch_2sm<EPI_STD, ACT_GELU_TANH>(grid, smem, stream, map_a, map_b, a, sc); break;
      case ACT_GELU_ERF: launch_2sm<EPI_STD, ACT_GELU_ERF>(grid, smem, stream, map_a, map_b, a, sc); break;
      case ACT_MISH: launch_2sm<EPI_STD, ACT_MISH>(grid, smem, stream, map_a, map_b, a, sc); break;
      default: fail("rowgemm_tc: unknown activation");
    }
  }
  B2_LAUNCH_CHECK();
  count_launch();
  return true;
}

}  // namespace

void rowgemm_tc(const RowGemm& p, const TcWeight& w, cudaStream_t stream) {
  B2_CHECK(w.ready, "rowgemm_tc: tensor-core weights not prepared");
  B2_CHECK(p.Cin == w.Cin && p.N == w.N && p.taps == w.taps && p.groups == w.groups, "rowgemm_tc: weight/problem mismatch");
  B2_CHECK((p.f16 != 0) == (w.f16 != 0), "rowgemm_tc: operand dtype differs from the prepared weight's (bf16 vs fp16)");
  B2_CHECK(p.N % 4 == 0 && p.ldo % 4 == 0 && p.o_shift % 4 == 0, "rowgemm_tc: output alignment");
  B2_CHECK(p.ldx % 8 == 0 && p.x_bstride % 8 == 0, "rowgemm_tc: A rows must be 16-byte aligned");
  B2_CHECK(p.groups == 1 || p.Cin % BK == 0, "rowgemm_tc: grouped problems need Cin % 64 == 0");
  B2_CHECK(p.M > 0 && p.B > 0, "rowgemm_tc: empty problem");
  const bool rope = p.rope_cs != nullptr;
  B2_CHECK(!rope || (p.rope_cols % 64 == 0 && p.vt_col0 % 32 == 0 && p.groups == 1 && p.B == 1),
           "rowgemm_tc: malformed rope epilogue");
  B2_CHECK(!rope || (p.gate == nullptr && p.res == nullptr && !p.accumulate && p.act == ACT_NONE && p.out2 == nullptr),
           "rowgemm_tc: the rope epilogue takes bias only");

  if (!p.accumulate && p.out2 == nullptr && p.o_shift == 0 && try_launch_2sm(p, w, stream, rope)) return;
  const int sms = sm_count();
  const int kchunks = ceil_div(p.Cin, BK);
  const TileShape ts = pick_tile(p.M, p.N, (long)p.B * p.groups, sms, rope);
  const int bm = 128 * ts.halves;
  CUtensorMap map_a, map_b;
  tc_encode_map(&map_a, p.x, (uint64_t)p.groups * p.Cin, (uint64_t)p.Lin, (uint64_t)p.B, (uint64_t)p.ldx,
                (uint64_t)p.x_bstride, (uint32_t)A_BOX_ROWS);
  tc_encode_map(&map_b, w.w.p, (uint64_t)w.Cin, (uint64_t)w.N, (uint64_t)w.groups * w.taps, (uint64_t)w.ldc,
                (uint64_t)w.N * w.ldc, (uint32_t)ts.bn);
  TcArgs a = make_args(p, ts.bn);
  Tc3Sched sc;
  sc.bm = bm; sc.halves = ts.halves;
  sc.m_tiles = ceil_div(p.M, bm);
  sc.n_tiles = ceil_div(p.N, ts.bn);
  const long tiles = (long)p.B * p.groups * sc.m_tiles * sc.n_tiles;
  B2_CHECK(tiles < (1L << 30), "rowgemm_tc: too many tiles");
  sc.num_tiles = (int)tiles;
  const int halo = (p.taps - 1) * p.dil;
  sc.a_rows = (int)round_up(bm + halo, A_BOX_ROWS);
  sc.a_box = sc.a_rows % 256 == 0 ? 256 : sc.a_rows % 128 == 0 ? 128 : 64;
  tc_encode_map(&map_a, p.x, (uint64_t)p.groups * p.Cin, (uint64_t)p.Lin, (uint64_t)p.B, (uint64_t)p.ldx,
                (uint64_t)p.x_bstride, (uint32_t)sc.a_box);
  sc.half_stride = half_stride_of(ts.bn);
  sc.nacc = (2 * ts.halves * sc.half_stride <= 512) ? 2 : 1;
  const int a_stage = sc.a_rows * 128, b_stage = ts.bn * 128;
  const int bar_bytes = (2 * MAX_A_STAGES + 2 * MAX_B_STAGES + 4) * 8 + 16;
  const int budget = 227 * 1024 - 1024 - bar_bytes - EPI_BYTES;
  // ring depths: a plain GEMM consumes one A and one B stage per 64-channel chunk -> equal depths; a convolution
  // consumes `taps` B stages per A stage -> two A stages, the rest of shared memory for B. A thin convolution whose whole
  // weight tensor fits beside two A stages keeps it resident (one fetch per CTA, no per-tap barrier traffic).
  int nA, nB;
  sc.bres = 0;
  const long w_bytes = (long)kchunks * p.taps * b_stage;
  if (p.taps > 1 && p.groups == 1 && sc.n_tiles == 1 && w_bytes + 2L * a_stage <= budget && tc_env().bres != 0) {
    sc.bres = 1;
    nB = kchunks * p.taps;
    nA = (int)((budget - w_bytes) / a_stage);
    if (nA > 4) nA = 4;
  } else if (p.taps == 1) {
    nA = budget / (a_stage + b_stage);
    if (nA > 6) nA = 6;
    if (nA < 1) nA = 1;
    nB = nA;
  } else {
    nA = 2;
    while (nA > 1 && budget - nA * a_stage < 2 * b_stage) --nA;
    nB = (budget - nA * a_stage) / b_stage;
    if (nB > MAX_B_STAGES) nB = MAX_B_STAGES;
  }
  B2_CHECK(nB >= 1, "rowgemm_tc: halo tile does not fit shared memory (kernel too long / dilation too large)");
  sc.nA = nA; sc.nB = nB;
  const int smem = nA * a_stage + nB * b_stage + EPI_BYTES + 1024 + bar_bytes;
  const int grid = sc.num_tiles < sms ? sc.num_tiles : sms;
  if (rope) {
    launch_kernel<EPI_ROPE, ACT_NONE>(grid, smem, stream, map_a, map_b, a, sc);
  } else {
    switch (p.act) {
      case ACT_NONE: launch_kernel<EPI_STD, ACT_NONE>(grid, smem, stream, map_a, map_b, a, sc); break;
      case ACT_GELU_TANH: launch_kernel<EPI_STD, ACT_GELU_TANH>(grid, smem, stream, map_a, map_b, a, sc); break;
      case ACT_GELU_ERF: launch_kernel<EPI_STD, ACT_GELU_ERF>(grid, smem, stream, map_a, map_b, a, sc); break;
      case ACT_MISH: launch_kernel<EPI_STD, ACT_MISH>(grid, smem, stream, map_a, map_b, a, sc); break;
      default: fail("rowgemm_tc: unknown activation");
    }
  }
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts

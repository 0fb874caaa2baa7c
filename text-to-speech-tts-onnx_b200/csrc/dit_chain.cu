// One DiT block's row-local chain as ONE persistent tcgen05 kernel (sm_100a).
//
// Why. For one utterance (R = 2 x 1126 rows) every dense layer of the DiT is a 5..14 GFLOP GEMM: as separate launches each
// one is a 15-25 us kernel of which 2-6 us are tensor work -- the rest is launch / fill / drain, and the LayerNorm-modulate
// between them is a kernel of its own (profiles/r01: GEMMs at 0.27 of the tensor peak, 1 395 rownorm launches per utterance).
// But everything between two attention calls is local to a block of rows:
//     att rows -> out-proj (+gate, +residual) -> LN-modulate -> ff1 (GELU) -> ff2 (+gate, +residual) -> LN-modulate -> q|k|v
// (modules.py:599-613; only attention itself mixes rows). So a TEAM of 8 CTA pairs owns a 256-row block and walks that
// chain without leaving the SMs: pair s of the team computes output columns [n/8 * s, n/8 * (s+1)) of every GEMM with
// tcgen05.mma.cta_group::2 (M = 256: each CTA holds 128 rows of A / D and half of the pair's weight slice), the activations
// between GEMMs go through L2 (written by the epilogue warps, fetched back by TMA), and the hand-offs inside a team are
// release / acquire counters in global memory -- no grid-wide barrier, no kernel boundary. 9 teams x 8 pairs = 144 CTAs for
// one utterance; for a batch the teams stride over the row blocks.
//
// Per CTA: 11 warps as in rowgemm_tc.cu (0-7 epilogue, 8 A producer, 9 MMA issuer (leader CTA only), 10 B producer).
//   job 0  out  : A = att16  K = D    pair columns 128  TMEM [0, 128)      epilogue: x += gate_msa * (acc + b)   + LN statistics
//   job 1  ff1  : A = n16    K = D    pair columns 256  TMEM [128, 384)    epilogue: GELU_tanh(acc + b) -> ff16
//   job 2  ff2  : A = ff16   K = FF   pair columns 128  TMEM [384, 512)    epilogue: x += gate_mlp * (acc + b)   + LN statistics
//   job 3  qkv  : A = n16b   K = D    pair columns 384  TMEM [128, 512)    epilogue: bias + interleaved RoPE, V transposed
// LayerNorm (no affine, eps 1e-6, modules.py:296) of a row needs all 1024 columns = all 8 pairs: each CTA reduces (sum, sum
// of squares) of its 128 columns in the epilogue it already runs, publishes them (8 partial pairs per row), waits for its
// team, then normalises + modulates its own 128 x 128 slab of x (re-read from L2) into the 16-bit A operand of the next GEMM.
// The weight stream never waits for anything (the B producer runs ahead across job boundaries and does not even wait for
// the previous kernel: weights are constants), so each hand-off costs the epilogue + one L2 round trip, not a pipeline refill.
#include "dit_chain.cuh"

#include <mutex>

#include "rowgemm_tc_dev.cuh"

namespace b200tts {

namespace {

constexpr int CH_A_STAGES = 4, CH_B_STAGES = 3;
constexpr int CH_A_BYTES = 128 * 128;            // 128 rows x 64 channels x 2 B
constexpr int CH_B_BYTES = 192 * 128;            // up to 192 weight rows per CTA and chunk (q|k|v: two 96-row boxes)
constexpr int CH_STAT_BYTES = 2 * 128 * 2 * 4;   // [column half e][row][sum, sum of squares]
constexpr int CH_EPI_BYTES = 8 * 8192;           // per epilogue warp: residual tile + result tile (4 KB each, rowgemm_tc_dev.cuh: EpiTile)
constexpr int CH_VEC_BYTES = 2 * 384 * 4;        // bias / gate of the CTA's column slice
constexpr int CH_TEAM = 8;                       // CTA pairs per row block
constexpr int CH_ROWS = 256;                     // rows per block (one M = 256 pair tile)
constexpr int CH_NFLAGS = 8;
enum { F_STAT1 = 0, F_N16 = 1, F_FF16 = 2, F_STAT2 = 3, F_N16B = 4 };
constexpr int CH_BAR_BYTES = (2 * CH_A_STAGES + 2 * CH_B_STAGES + 8 + 8) * 8 + 16;
constexpr int CH_SMEM = CH_A_STAGES * CH_A_BYTES + CH_B_STAGES * CH_B_BYTES + CH_EPI_BYTES + CH_VEC_BYTES + CH_STAT_BYTES + CH_BAR_BYTES + 1024;
static_assert(CH_SMEM <= 227 * 1024, "dit_chain: shared memory budget");

struct ChainArgs {
  int R, nrb, teams, has_qkv, f16, D, FF;
  float* x;
  __nv_bfloat16 *n16, *ff16, *n16b;
  const float *b_out, *gate_msa, *shift_mlp, *scale_mlp, *b_ff1, *b_ff2, *gate_mlp, *shift_nxt, *scale_nxt, *b_qkv;
  __nv_bfloat16* qk16;
  const __half2* rope_cs;
  int rope_rows;
  __nv_bfloat16* vt_out;
  int vt_ld, vt_heads;
  float* stats;
  unsigned* flags;
  unsigned long long* trace;      // optional [CTA][64] globaltimer stamps (B200TTS_CHAIN_TRACE, tools/chain_trace.py)
};

// ---- team hand-offs: monotonic counters in global memory (zero at kernel start), one per (row block, event) -----------------
// Polling is RELAXED (an acquire load costs a CCTL.IVALL -- a full L1 invalidation of the SM -- per poll: ncu r02a counted
// 92 865 of them in one launch, thrashing the L1 lines the epilogue warps of the same SM live on); one acquire fence follows
// the poll that succeeds.
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_relaxed_gpu(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy global writes (epilogue stores) <-> async-proxy reads (TMA loads of the same bytes by another CTA)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// Spin until the counter reaches `target`. A team never waits on anything but its own members, which are co-resident by
// construction (dit_chain(): grid <= resident pairs); the watchdog turns a protocol bug into a trap instead of a hung GPU.
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
  const long long t0 = clock64();
  while (ld_relaxed_gpu(p) < target) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
  fence_acq_rel_gpu();
}
__device__ __forceinline__ void stamp(const ChainArgs& c, int slot) {
  if (c.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    c.trace[(size_t)blockIdx.x * 64 + slot] = t;
  }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// The eight epilogue warps of a CTA have finished writing a slab: make it visible to the team and count this CTA in. The CTA
// barrier orders every warp's stores before thread 0's fence, and a gpu-scope fence is cumulative over what its thread has
// observed (the grid-barrier idiom: bar.sync; thread 0: fence + atomic) -- one MEMBAR per CTA instead of 256.
__device__ __forceinline__ void team_signal(unsigned* flag, int tid) {
  epi_bar();
  if (tid == 0) {
    fence_proxy_async_global();
    fence_acq_rel_gpu();
    red_relaxed_gpu(flag, 1u);
  }
}

struct JobShape { int K, n_pair, nsub, sub_n, tmem_col, b_rows; };
__device__ __forceinline__ JobShape job_shape(int j, int D, int FF) {
  switch (j) {
    case 0: return JobShape{D, D / CH_TEAM, 1, D / CH_TEAM, 0, D / (2 * CH_TEAM)};
    case 1: return JobShape{D, FF / CH_TEAM, 1, FF / CH_TEAM, 128, FF / (2 * CH_TEAM)};
    case 2: return JobShape{FF, D / CH_TEAM, 1, D / CH_TEAM, 384, D / (2 * CH_TEAM)};
    default: return JobShape{D, 3 * D / CH_TEAM, 2, 3 * D / (2 * CH_TEAM), 128, 3 * D / (4 * CH_TEAM)};
  }
}

// LayerNorm statistics + modulation of this CTA's [128 rows] x [128 columns at col0] slab of x (all 256 epilogue threads).
//   rsum / rsq : this lane's row sums over the warp's column blocks (epilogue_rows_tma<.., STATS>)
__device__ __forceinline__ void ln_phase(const ChainArgs& c, int rb, int rank, int slice, int kind, float rsum, float rsq,
                                         const float* scale, const float* shift, __nv_bfloat16* dst, unsigned* flag_stat,
                                         unsigned* flag_ready, float* st, int warp, int lane) {
  const int tid = warp * 32 + lane;
  const int q = warp & 3, e = warp >> 2;
  st[(e * 128 + q * 32 + lane) * 2 + 0] = rsum;                 // the row-layout epilogue leaves a row's sums in one lane
  st[(e * 128 + q * 32 + lane) * 2 + 1] = rsq;
  epi_bar();
  float* stats = c.stats + ((size_t)(rb * 2 + kind) * CH_ROWS + rank * 128) * (CH_TEAM * 2);
  if (tid < 128) {
    const float s = st[tid * 2] + st[(128 + tid) * 2], sq = st[tid * 2 + 1] + st[(128 + tid) * 2 + 1];
    *reinterpret_cast<float2*>(stats + ((size_t)tid * CH_TEAM + slice) * 2) = make_float2(s, sq);
  }
  epi_bar();
  if (tid == 0) {
    fence_acq_rel_gpu();
    red_relaxed_gpu(flag_stat, 1u);
    wait_counter(flag_stat, 2u * CH_TEAM);
  }
  epi_bar();
  if (tid == 0) stamp(c, 8 + 8 * (kind * 2) + 6);
  // pass 2: warp w normalises rows [16w, 16w + 16) of the slab; lane l owns columns col0 + 4l .. 4l + 3. All loads of the
  // 16 rows are issued before the first use (one L2 round trip for the statistics, one for x, not one per row).
  const int col = slice * (c.D / CH_TEAM) + lane * 4;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + col));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + col));
  const float inv_d = 1.0f / (float)c.D;
  const long grow0 = (long)rb * CH_ROWS + rank * 128 + warp * 16;
  float2 pr[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)       // lane = (row k*4 + lane/8, slot lane%8)
    pr[k] = __ldcg(reinterpret_cast<const float2*>(stats + ((size_t)(warp * 16 + k * 4 + (lane >> 3)) * CH_TEAM + (lane & 7)) * 2));
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float4 xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long g = grow0 + half * 8 + j;
      xv[j] = g < c.R ? __ldcg(reinterpret_cast<const float4*>(c.x + g * c.D + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (half == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          pr[k].x += __shfl_xor_sync(0xffffffffu, pr[k].x, o);
          pr[k].y += __shfl_xor_sync(0xffffffffu, pr[k].y, o);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int jj = half * 8 + j;
      const float px = (jj >> 2) == 0 ? pr[0].x : (jj >> 2) == 1 ? pr[1].x : (jj >> 2) == 2 ? pr[2].x : pr[3].x;
      const float py = (jj >> 2) == 0 ? pr[0].y : (jj >> 2) == 1 ? pr[1].y : (jj >> 2) == 2 ? pr[2].y : pr[3].y;
      const float sum = __shfl_sync(0xffffffffu, px, (jj & 3) * 8), sq = __shfl_sync(0xffffffffu, py, (jj & 3) * 8);
      const long g = grow0 + jj;
      if (g >= c.R) continue;                                   // warp-uniform
      const float mean = sum * inv_d;
      const float var = fmaxf(sq * inv_d - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-6f);
      const float4 v = xv[j];
      const float y0 = (v.x - mean) * rstd * (1.0f + sc.x) + sh.x, y1 = (v.y - mean) * rstd * (1.0f + sc.y) + sh.y;
      const float y2 = (v.z - mean) * rstd * (1.0f + sc.z) + sh.z, y3 = (v.w - mean) * rstd * (1.0f + sc.w) + sh.w;
      *reinterpret_cast<uint2*>(dst + g * c.D + col) = pack16x4(y0, y1, y2, y3, c.f16);
    }
  }
  team_signal(flag_ready, tid);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS3, 1)
dit_chain_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
                 const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mA3,
                 const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1,
                 const __grid_constant__ CUtensorMap mB2, const __grid_constant__ CUtensorMap mB3,
                 const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mFFo,
                 const __grid_constant__ CUtensorMap mQKo, const ChainArgs c) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + CH_A_STAGES * CH_A_BYTES;
  uint8_t* smem_epi = smem_b + CH_B_STAGES * CH_B_BYTES;              // 1024-aligned: stage sizes are multiples of 1 KB
  float* smem_vec = reinterpret_cast<float*>(smem_epi + CH_EPI_BYTES);
  float* smem_stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem_vec) + CH_VEC_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_stat) + CH_STAT_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + CH_A_STAGES;
  uint64_t* b_full = a_empty + CH_A_STAGES;
  uint64_t* b_empty = b_full + CH_B_STAGES;
  uint64_t* acc_full = b_empty + CH_B_STAGES;            // [4] one per job
  uint64_t* acc_empty = acc_full + 4;                    // [4]
  uint64_t* res_bar = acc_empty + 4;                     // [8] one per epilogue warp (residual tiles)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int team = pair / CH_TEAM, slice = pair - team * CH_TEAM;
  const int njobs = c.has_qkv ? 4 : 3;

  if (warp == WARP_TMA && lane == 0) {
    prefetch_tmap(&mA0); prefetch_tmap(&mA1); prefetch_tmap(&mA2); prefetch_tmap(&mA3);
    prefetch_tmap(&mB0); prefetch_tmap(&mB1); prefetch_tmap(&mB2); prefetch_tmap(&mB3);
    prefetch_tmap(&mX); prefetch_tmap(&mFFo); prefetch_tmap(&mQKo);
    for (int s = 0; s < CH_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < CH_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int j = 0; j < 4; ++j) { mbar_init(&acc_full[j], 1); mbar_init(&acc_empty[j], 16); }
    for (int w = 0; w < 8; ++w) mbar_init(&res_bar[w], 1);
    fence_barrier_init();
  }
  cluster_sync_all();                                     // the peer's barriers exist before anything signals them
  if (warp == WARP_MMA) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  cluster_sync_all();                                     // both allocations are done before the leader's MMAs write the peer's TMEM
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      // ===== A producer (both CTAs): this CTA's 128 rows of each job's A operand; waits for the team hand-off first =====
      pdl_wait();
      int sa = 0; uint32_t pa = 0;
      for (int rb = team; rb < c.nrb; rb += c.teams) {
        const int row0 = rb * CH_ROWS + (int)rank * 128;
        const unsigned* flags = c.flags + (size_t)rb * CH_NFLAGS;
        for (int j = 0; j < njobs; ++j) {
          const JobShape js = job_shape(j, c.D, c.FF);
          const CUtensorMap* map = j == 0 ? &mA0 : j == 1 ? &mA1 : j == 2 ? &mA2 : &mA3;
          stamp(c, 8 + 8 * j + 0);
          if (j > 0) {
            wait_counter(flags + (j == 1 ? F_N16 : j == 2 ? F_FF16 : F_N16B), 2u * CH_TEAM);
            fence_proxy_async_global();
          }
          stamp(c, 8 + 8 * j + 1);
          for (int ck = 0; ck < js.K / BK; ++ck) {
            mbar_wait(&a_empty[sa], pa ^ 1);
            if (leader) mbar_expect_tx(&a_full[sa], 2u * CH_A_BYTES);
            tma_load_3d_2sm(smem_a + sa * CH_A_BYTES, map, &a_full[sa], ck * BK, row0, 0);
            if (++sa == CH_A_STAGES) { sa = 0; pa ^= 1; }
          }
        }
      }
    }
  } else if (warp == WARP_TMA_B) {
    if (lane == 0) {
      // ===== B producer (both CTAs): this CTA's half of the pair's weight slice. No pdl_wait: weights are constants =====
      int sb = 0; uint32_t pb = 0;
      for (int rb = team; rb < c.nrb; rb += c.teams) {
        for (int j = 0; j < njobs; ++j) {
          const JobShape js = job_shape(j, c.D, c.FF);
          const CUtensorMap* map = j == 0 ? &mB0 : j == 1 ? &mB1 : j == 2 ? &mB2 : &mB3;
          const int n_row0 = slice * js.n_pair + (int)rank * js.b_rows;
          const uint32_t tx = 2u * (uint32_t)(js.nsub * js.b_rows * 128);
          for (int ck = 0; ck < js.K / BK; ++ck) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            if (leader) mbar_expect_tx(&b_full[sb], tx);
            for (int s = 0; s < js.nsub; ++s)
              tma_load_3d_2sm(smem_b + sb * CH_B_BYTES + s * (CH_B_BYTES / 2), map, &b_full[sb], ck * BK, n_row0 + s * js.sub_n, 0);
            if (++sb == CH_B_STAGES) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    if (leader) {
      // ===== MMA issuer (leader CTA only): M = 256 across the pair =====
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a)), b_lo0 = desc_lo_sw128(smem_u32(smem_b));
      const uint32_t a_stage_lo = (uint32_t)CH_A_BYTES >> 4, b_stage_lo = (uint32_t)CH_B_BYTES >> 4, b_sub_lo = (uint32_t)(CH_B_BYTES / 2) >> 4;
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int rb = team; rb < c.nrb; rb += c.teams, ++it) {
        for (int j = 0; j < njobs; ++j) {
          const JobShape js = job_shape(j, c.D, c.FF);
          const uint32_t idesc = idesc_f16kind(256, js.sub_n, c.f16);
          mbar_wait(&acc_empty[j], ((uint32_t)it & 1u) ^ 1u);
          // the q|k|v accumulator [128, 512) overlaps ff1's and ff2's: the previous row block's q|k|v epilogue must have drained
          if (j == 1 && it > 0 && c.has_qkv) mbar_wait(&acc_empty[3], (uint32_t)(it - 1) & 1u);
          tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)js.tmem_col;
          const int chunks = js.K / BK;
          for (int ck = 0; ck < chunks; ++ck) {
            mbar_wait(&a_full[sa], pa);
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            if (ck == 0 && lane == 0) stamp(c, 8 + 8 * j + 2);
            const uint32_t a_lo = a_lo0 + (uint32_t)sa * a_stage_lo;
            const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_stage_lo;
            const uint32_t accum = ck > 0 ? 1u : 0u;
            if (elect_one()) {
              if (js.nsub == 1) {
                umma2_bf16_lohi(d0, a_lo + 0, b_lo + 0, idesc, accum);
                umma2_bf16_lohi(d0, a_lo + 2, b_lo + 2, idesc, 1u);
                umma2_bf16_lohi(d0, a_lo + 4, b_lo + 4, idesc, 1u);
                umma2_bf16_lohi(d0, a_lo + 6, b_lo + 6, idesc, 1u);
              } else {                                      // two column halves share the A tile
                const uint32_t d1 = d0 + (uint32_t)js.sub_n, b_hi = b_lo + b_sub_lo;
                umma2_bf16_lohi(d0, a_lo + 0, b_lo + 0, idesc, accum);
                umma2_bf16_lohi(d1, a_lo + 0, b_hi + 0, idesc, accum);
                umma2_bf16_lohi(d0, a_lo + 2, b_lo + 2, idesc, 1u);
                umma2_bf16_lohi(d1, a_lo + 2, b_hi + 2, idesc, 1u);
                umma2_bf16_lohi(d0, a_lo + 4, b_lo + 4, idesc, 1u);
                umma2_bf16_lohi(d1, a_lo + 4, b_hi + 4, idesc, 1u);
                umma2_bf16_lohi(d0, a_lo + 6, b_lo + 6, idesc, 1u);
                umma2_bf16_lohi(d1, a_lo + 6, b_hi + 6, idesc, 1u);
              }
              umma2_commit_mc(&b_empty[sb]);
              umma2_commit_mc(&a_empty[sa]);
              if (ck == chunks - 1) { umma2_commit_mc(&acc_full[j]); stamp(c, 8 + 8 * j + 3); }
            }
            __syncwarp();
            if (++sa == CH_A_STAGES) { sa = 0; pa ^= 1; }
            if (++sb == CH_B_STAGES) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs, warps 0..7): this CTA's 128 rows; q = TMEM lane quarter, e = even / odd 32-column blocks.
    // Row-layout epilogue with TMA tiles (rowgemm_tc_dev.cuh: epilogue_rows_tma) =====
    const int q = warp & 3, e = warp >> 2;
    const int tid = warp * 32 + lane;
    if (tid == 0) stamp(c, 0);
    pdl_wait();
    if (tid == 0) stamp(c, 1);
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    EpiTile et;
    et.res_tile = smem_epi + warp * 8192; et.out_tile = et.res_tile + 4096; et.res_bar = &res_bar[warp]; et.res_phase = 0u;
    float* s_bias = smem_vec;
    float* s_gate = smem_vec + 384;
    // bias / gate of the CTA's column slice -> shared memory (broadcast reads in the epilogue)
    auto stage_vec = [&](const float* bias, const float* gate, int n0, int n) {
      epi_bar();                                               // the previous job's readers are done
      for (int i = tid; i < n; i += 256) {
        s_bias[i] = __ldg(bias + n0 + i);
        if (gate) s_gate[i] = __ldg(gate + n0 + i);
      }
      epi_bar();
    };
    TcArgs a;
    a.taps = 1; a.dil = 1; a.center = 0; a.groups = 1; a.M = c.R;
    a.o_bstride = 0; a.o_shift = 0; a.accumulate = 0; a.scale = 1.0f; a.out2 = nullptr; a.f16 = c.f16;
    a.rope_cs = nullptr; a.rope_cols = 0; a.rope_rows = 1; a.vt_out = nullptr; a.vt_col0 = 0; a.vt_ld = 0; a.vt_heads = 0;
    a.out = nullptr; a.ldo = 0; a.o_limit = 0; a.out_bf16 = 0; a.bias = nullptr; a.gate = nullptr; a.res = nullptr;
    int it = 0;
    for (int rb = team; rb < c.nrb; rb += c.teams, ++it) {
      unsigned* flags = c.flags + (size_t)rb * CH_NFLAGS;
      const uint32_t ph = (uint32_t)it & 1u;
      const int row0 = rb * CH_ROWS + (int)rank * 128 + q * 32;
      float rsum, rsq;
      // ---- job 0: x += gate_msa * (att @ Wout^T + b_out) ; LN-modulate (mlp) -> n16 ----
      {
        const JobShape js = job_shape(0, c.D, c.FF);
        const int n0 = slice * js.n_pair;
        a.Cin = js.K; a.N = c.D; a.BN = js.n_pair; a.kchunks = js.K / BK;
        stage_vec(c.b_out, c.gate_msa, n0, js.n_pair);
        epi_tma_fetch_res(&mX, et, n0 + e * 32, row0, lane);
        rsum = 0.f; rsq = 0.f;
        mbar_wait(&acc_full[0], ph);
        tc_fence_after();
        if (tid == 0) stamp(c, 8 + 8 * 0 + 4);
        epilogue_rows_tma<TK_RES_F32, ACT_NONE, true>(a, &mX, &mX, et, tmem_base + (uint32_t)js.tmem_col + lane_sel, row0, n0, e * 32, 64, lane,
                                                      s_bias, s_gate, rsum, rsq);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&acc_empty[0]);
        epi_tma_drain(lane);
        if (tid == 0) stamp(c, 8 + 8 * 0 + 5);
        ln_phase(c, rb, (int)rank, slice, 0, rsum, rsq, c.scale_mlp, c.shift_mlp, c.n16, flags + F_STAT1, flags + F_N16, smem_stat, warp, lane);
        if (tid == 0) stamp(c, 8 + 8 * 0 + 7);
      }
      // ---- job 1: ff16 = GELU_tanh(n16 @ Wff1^T + b_ff1) ----
      {
        const JobShape js = job_shape(1, c.D, c.FF);
        const int n0 = slice * js.n_pair;
        a.Cin = js.K; a.N = c.FF; a.BN = js.n_pair; a.kchunks = js.K / BK;
        stage_vec(c.b_ff1, nullptr, n0, js.n_pair);
        mbar_wait(&acc_full[1], ph);
        tc_fence_after();
        if (tid == 0) stamp(c, 8 + 8 * 1 + 4);
        epilogue_rows_tma<TK_ACT16, ACT_GELU_TANH, false>(a, &mFFo, nullptr, et, tmem_base + (uint32_t)js.tmem_col + lane_sel, row0, n0, e * 32, 64,
                                                          lane, s_bias, nullptr, rsum, rsq);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&acc_empty[1]);
        epi_tma_drain(lane);
        if (tid == 0) stamp(c, 8 + 8 * 1 + 5);
        team_signal(flags + F_FF16, tid);
        if (tid == 0) stamp(c, 8 + 8 * 1 + 7);
      }
      // ---- job 2: x += gate_mlp * (ff16 @ Wff2^T + b_ff2) ; LN-modulate (next block's attention / final) -> n16b ----
      {
        const JobShape js = job_shape(2, c.D, c.FF);
        const int n0 = slice * js.n_pair;
        a.Cin = js.K; a.N = c.D; a.BN = js.n_pair; a.kchunks = js.K / BK;
        stage_vec(c.b_ff2, c.gate_mlp, n0, js.n_pair);
        epi_tma_fetch_res(&mX, et, n0 + e * 32, row0, lane);
        rsum = 0.f; rsq = 0.f;
        mbar_wait(&acc_full[2], ph);
        tc_fence_after();
        if (tid == 0) stamp(c, 8 + 8 * 2 + 4);
        epilogue_rows_tma<TK_RES_F32, ACT_NONE, true>(a, &mX, &mX, et, tmem_base + (uint32_t)js.tmem_col + lane_sel, row0, n0, e * 32, 64, lane,
                                                      s_bias, s_gate, rsum, rsq);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&acc_empty[2]);
        epi_tma_drain(lane);
        if (tid == 0) stamp(c, 8 + 8 * 2 + 5);
        ln_phase(c, rb, (int)rank, slice, 1, rsum, rsq, c.scale_nxt, c.shift_nxt, c.n16b, flags + F_STAT2, flags + F_N16B, smem_stat, warp, lane);
        if (tid == 0) stamp(c, 8 + 8 * 2 + 7);
      }
      // ---- job 3: q | k | v of the next block: bias + RoPE -> qk16 (TMA tiles), V transposed -> vt_out ----
      if (c.has_qkv) {
        const JobShape js = job_shape(3, c.D, c.FF);
        const int n0 = slice * js.n_pair;
        a.Cin = js.K; a.N = 3 * c.D; a.BN = js.n_pair; a.kchunks = js.K / BK;
        a.rope_cs = c.rope_cs; a.rope_cols = 2 * c.D; a.rope_rows = c.rope_rows;
        a.vt_out = c.vt_out; a.vt_col0 = 2 * c.D; a.vt_ld = c.vt_ld; a.vt_heads = c.vt_heads;
        stage_vec(c.b_qkv, nullptr, n0, js.n_pair);
        mbar_wait(&acc_full[3], ph);
        tc_fence_after();
        if (tid == 0) stamp(c, 8 + 8 * 3 + 4);
        epilogue_rows_tma<TK_ROPE16, ACT_NONE, false>(a, &mQKo, nullptr, et, tmem_base + (uint32_t)js.tmem_col + lane_sel, row0, n0, e * 32, 64, lane,
                                                      s_bias, nullptr, rsum, rsq);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&acc_empty[3]);
        epi_tma_drain(lane);
        if (tid == 0) stamp(c, 8 + 8 * 3 + 5);
        a.rope_cs = nullptr; a.rope_cols = 0; a.rope_rows = 1; a.vt_out = nullptr; a.vt_col0 = 0;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// CTA pairs that can be resident at once: one CTA per SM (shared memory), and clusters of two pack the TPCs exactly.
int resident_pairs() {
  static int pairs = -1;
  static std::once_flag once;
  std::call_once(once, [] {
    B2_CUDA(cudaFuncSetAttribute(dit_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
    int dev = 0, sms = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    pairs = sms / 2;
  });
  return pairs;
}

}  // namespace

size_t dit_chain_stats_floats(int R) { return (size_t)ceil_div(R, CH_ROWS) * 2 * CH_ROWS * CH_TEAM * 2; }
size_t dit_chain_flag_words(int R) { return (size_t)ceil_div(R, CH_ROWS) * CH_NFLAGS; }

bool dit_chain_supported(int D, int FF, int H) {
  // the job table (column slices of 128 / 256 / 128 / 2 x 192 per pair, 512 TMEM columns) is laid out for the F5 DiT
  if (D != 1024 || FF != 2048 || H * 64 != D) return false;
  return resident_pairs() >= CH_TEAM;
}

void dit_chain(const DitChain& d, cudaStream_t stream) {
  B2_CHECK(dit_chain_supported(d.D, d.FF, d.D / 64), "dit_chain: unsupported shape or device");
  B2_CHECK(d.R > 0 && d.att16 && d.x && d.n16 && d.ff16 && d.n16b && d.stats && d.flags, "dit_chain: null buffer");
  B2_CHECK(d.w_out && d.w_ff1 && d.w_ff2 && d.w_out->ready && d.w_ff1->ready && d.w_ff2->ready, "dit_chain: weights not prepared");
  B2_CHECK(!d.has_qkv || (d.w_qkv && d.w_qkv->ready && d.qk16 && d.rope_cs && d.vt_out), "dit_chain: q|k|v phase arguments");
  const TcWeight* wq = d.has_qkv ? d.w_qkv : d.w_out;       // (a valid map for the unused slot)
  for (const TcWeight* w : {d.w_out, d.w_ff1, d.w_ff2, wq})
    B2_CHECK((w->f16 != 0) == (d.f16 != 0) && w->taps == 1 && w->groups == 1, "dit_chain: weight dtype / layout");
  B2_CHECK(d.w_out->N == d.D && d.w_out->Cin == d.D && d.w_ff1->N == d.FF && d.w_ff1->Cin == d.D && d.w_ff2->N == d.D &&
           d.w_ff2->Cin == d.FF && (!d.has_qkv || (d.w_qkv->N == 3 * d.D && d.w_qkv->Cin == d.D)), "dit_chain: weight shapes");
  ChainArgs c;
  c.R = d.R; c.nrb = ceil_div(d.R, CH_ROWS); c.has_qkv = d.has_qkv; c.f16 = d.f16; c.D = d.D; c.FF = d.FF;
  const int max_teams = resident_pairs() / CH_TEAM;
  c.teams = c.nrb < max_teams ? c.nrb : max_teams;
  c.x = d.x; c.n16 = d.n16; c.ff16 = d.ff16; c.n16b = d.n16b;
  c.b_out = d.b_out; c.gate_msa = d.gate_msa; c.shift_mlp = d.shift_mlp; c.scale_mlp = d.scale_mlp; c.b_ff1 = d.b_ff1; c.b_ff2 = d.b_ff2;
  c.gate_mlp = d.gate_mlp; c.shift_nxt = d.shift_nxt; c.scale_nxt = d.scale_nxt; c.b_qkv = d.b_qkv;
  c.qk16 = d.qk16; c.rope_cs = d.rope_cs; c.rope_rows = d.rope_rows > 0 ? d.rope_rows : 1; c.vt_out = d.vt_out; c.vt_ld = d.vt_ld; c.vt_heads = d.vt_heads;
  c.stats = d.stats; c.flags = d.flags; c.trace = d.trace;
  CUtensorMap mA[4], mB[4];
  const void* a_ptr[4] = {d.att16, d.n16, d.ff16, d.n16b};
  const int a_k[4] = {d.D, d.D, d.FF, d.D};
  for (int j = 0; j < 4; ++j)
    tc_encode_map(&mA[j], a_ptr[j], (uint64_t)a_k[j], (uint64_t)d.R, 1, (uint64_t)a_k[j], (uint64_t)d.R * a_k[j], 128);
  const TcWeight* ws[4] = {d.w_out, d.w_ff1, d.w_ff2, wq};
  const uint32_t b_box[4] = {(uint32_t)(d.D / (2 * CH_TEAM)), (uint32_t)(d.FF / (2 * CH_TEAM)), (uint32_t)(d.D / (2 * CH_TEAM)),
                             d.has_qkv ? (uint32_t)(3 * d.D / (4 * CH_TEAM)) : (uint32_t)(d.D / (2 * CH_TEAM))};
  for (int j = 0; j < 4; ++j)
    tc_encode_map(&mB[j], ws[j]->w.p, (uint64_t)ws[j]->Cin, (uint64_t)ws[j]->N, 1, (uint64_t)ws[j]->ldc, (uint64_t)ws[j]->N * ws[j]->ldc, b_box[j]);
  // epilogue tiles: x (fp32, residual in / result out), ff16 and q|k (16 bit, result out), 32 rows x 32 columns each
  CUtensorMap mX, mFFo, mQKo;
  tc_encode_map2d(&mX, d.x, 4, (uint64_t)d.D, (uint64_t)d.R, (uint64_t)d.D, 32, 32);
  tc_encode_map2d(&mFFo, d.ff16, 2, (uint64_t)d.FF, (uint64_t)d.R, (uint64_t)d.FF, 32, 32);
  if (d.has_qkv) tc_encode_map2d(&mQKo, d.qk16, 2, (uint64_t)2 * d.D, (uint64_t)d.R, (uint64_t)2 * d.D, 32, 32);
  else mQKo = mFFo;
  launch_pdl(dit_chain_kernel, dim3((unsigned)(c.teams * CH_TEAM * 2)), dim3(NTHREADS3), (size_t)CH_SMEM, stream, mA[0], mA[1], mA[2], mA[3],
             mB[0], mB[1], mB[2], mB[3], mX, mFFo, mQKo, c);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts

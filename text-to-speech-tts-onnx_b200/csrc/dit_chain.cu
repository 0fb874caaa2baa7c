// One DiT block's row-local chain as ONE persistent tcgen05 kernel (sm_100a).
//
// Why. Everything between two attention calls is local to a block of rows:
//     att rows -> out-proj (+gate, +residual) -> LN-modulate -> ff1 (GELU) -> ff2 (+gate, +residual) -> LN-modulate -> q|k|v
// (modules.py:599-613; only attention itself mixes rows). As separate launches each of these is a kernel whose launch, fill,
// drain and fp32-residual epilogue are exposed, with a LayerNorm kernel between them (profiles/r01: GEMMs at 0.27 of the tensor
// peak for one utterance, 1 395 rownorm launches). Here a TEAM of T CTA pairs owns a 256-row block and walks that chain without
// leaving the SMs: pair s of the team computes output columns [n/T * s, n/T * (s+1)) of every GEMM with
// tcgen05.mma.cta_group::2 (M = 256: each CTA holds 128 rows of A / D and half of the pair's weight slice), in sub-tiles of
// <= 256 columns whose accumulators alternate between the two halves of TMEM (the epilogue of one sub-tile runs under the main
// loop of the next). The activations between GEMMs go through L2 (written by the epilogue warps, fetched back by TMA), and the
// hand-offs inside a team are release / acquire counters in global memory -- no grid-wide barrier, no kernel boundary.
//   T = 8 for one utterance (9 row blocks x 8 pairs = 144 CTAs: every GEMM is one sub-tile per pair),
//   T = 1 for a batch of 8 (71 row blocks, one pair each: 4 / 8 / 4 / 12 sub-tiles per GEMM, hand-offs stay inside the pair).
// A launch runs in two phases (dit_chain_plan): whole rounds of row blocks at one team size, then the remaining blocks with a
// larger team each, so that a block count just above a multiple of the resident pairs does not cost a whole extra round.
// Optional e4m3 operands (template FP8 = 1: ff1 and q|k|v, 2: ff2 as well) on tcgen05 kind::f8f6f4; see DESIGN.md 3a item 8.
//
// Per CTA: 11 warps as in rowgemm_tc.cu (0-7 epilogue, 8 A producer, 9 MMA issuer (leader CTA only), 10 B producer).
//   job 0  out  : A = att16  K = D    epilogue: x += gate_msa * (acc + b); n16 = c x (1 + scale_mlp); LN partials
//   job 1  ff1  : A = n16    K = D    epilogue: GELU_tanh(rho acc - rmu u + v) -> ff16
//   job 2  ff2  : A = ff16   K = FF   epilogue: x += gate_mlp * (acc + b); n16b = c x (1 + scale_nxt); LN partials
//   job 3  qkv  : A = n16b   K = D    epilogue: rho acc - rmu u + v, interleaved RoPE, V transposed   (the NEXT block's projections)
// LayerNorm (no affine, eps 1e-6, modules.py:296) of a row needs all 1024 columns = all T pairs, i.e. it can only be applied
// after the whole row has been produced. Until round 2b that was a second pass: publish (sum, sum of squares), wait for the
// team, re-read the CTA's slab of x from L2, normalise + modulate, write the 16-bit A operand (14 of 57 us per block for one
// utterance, 2 x 28 us of a 300 us launch for eight: profiles/r02). Now the LayerNorm is FOLDED INTO THE GEMMS around it:
//     (LN(x)(1 + s) + t) W^T + b  =  rstd * (x (1 + s)) W^T  -  rstd * mean * u  +  v,    u = W (1 + s),  v = W t + b
// (u, v: per-column vectors of a (block, Euler step), computed once per operand type from the 16-bit weights, f5.cu). The
// epilogue that produces x writes  c * x * (1 + s)  as the 16-bit A operand in the same pass (plain 64-byte stores per lane)
// and publishes the row partials; the NEXT GEMM's epilogue computes  (rstd / c) * acc - rstd * mean * u + v. The statistics
// are needed only there, a whole main loop later, so no CTA ever waits for them. c is the row's 1 / std at the previous
// LayerNorm (carried in `rowscale`): algebraically irrelevant, it keeps the fp16 operand in the range normalised rows have.
// Only the last block's final LayerNorm (it feeds proj_out, a separate GEMM) still takes the second pass (ln_phase).
// The weight stream never waits for anything (the B producer runs ahead across job boundaries and does not even wait for the
// previous kernel: weights are constants).
#include "dit_chain.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "rowgemm_tc_dev.cuh"

namespace b200tts {

namespace {

constexpr int CH_A_STAGES = 4, CH_B_STAGES = 4;
constexpr int CH_A_BYTES = 128 * 128;            // 128 rows x 64 channels x 2 B
constexpr int CH_B_BYTES = 128 * 128;            // up to 128 weight rows per CTA and chunk (half of a 256-column sub-tile)
constexpr int CH_STAT_BYTES = 8 * 2 * 128 * 2 * 4;   // [128-column group of the slice][column parity e][row][sum, sum of squares]
constexpr int CH_EPI_BYTES = 8 * 8192;           // per epilogue warp: residual tile + result tile (4 KB each, rowgemm_tc_dev.cuh: EpiTile)
constexpr int CH_VEC_BYTES = 3 * 256 * 4;        // bias / gate (or u) / 1 + scale of the current sub-tile
constexpr int CH_MAX_TEAM = 8;                   // statistics slots per row
constexpr int CH_ROWS = 256;                     // rows per block (one M = 256 pair tile)
constexpr int CH_NFLAGS = 8;
enum { F_STAT1 = 0, F_N16 = 1, F_FF16 = 2, F_STAT2 = 3, F_N16B = 4 };
constexpr int CH_BAR_BYTES = (2 * CH_A_STAGES + 2 * CH_B_STAGES + 4 + 8) * 8 + 16;
constexpr int CH_SMEM = CH_A_STAGES * CH_A_BYTES + CH_B_STAGES * CH_B_BYTES + CH_EPI_BYTES + CH_VEC_BYTES + CH_STAT_BYTES + CH_BAR_BYTES + 1024;
static_assert(CH_SMEM <= 227 * 1024, "dit_chain: shared memory budget");

struct ChainArgs {
  int R, nrb, teams, team, has_qkv, f16, D, FF;      // phase 0: `teams` teams of `team` CTA pairs (1, 2, 4 or 8) walk row blocks [0, nrb0)
  int nrb0, rem, team1;                              // phase 1: the `rem` = nrb - nrb0 remaining row blocks, one team of `team1` pairs each
  int l2hint;                                        // experiment: x tiles stored with the L2 evict_last policy
  float* x;
  __nv_bfloat16 *n16, *ff16, *n16b;
  const float *b_out, *gate_msa, *shift_mlp, *scale_mlp, *b_ff1, *b_ff2, *gate_mlp, *shift_nxt, *scale_nxt, *b_qkv;
  __nv_bfloat16* qk16;
  const __half2* rope_cs;
  int rope_rows;
  const int2* rowinfo;
  __nv_bfloat16* vt_out;
  int vt_ld, vt_heads;
  // LayerNorm folded into the GEMMs (see the header): per-column vectors of this (block, Euler step) and the per-row scale
  const float *u_ff1, *v_ff1, *u_qkv, *v_qkv;
  float* rowscale;
  int fp8;                                           // ff1 and q|k|v run with e4m3 operands (A emitted as e4m3, weights pre-quantised)
  const float *sw_ff1, *sw_qkv;                      // per-output-channel scales of the e4m3 weights
  float* stats;
  unsigned* flags;
  unsigned long long* trace;      // optional [CTA][64] globaltimer stamps (B200TTS_CHAIN_TRACE, tools/chain_trace.py)
};

// ---- team hand-offs: monotonic counters in global memory (zero at kernel start), one per (row block, event) -----------------
// Polling is RELAXED (an acquire load costs a CCTL.IVALL -- a full L1 invalidation of the SM -- per poll: ncu r02a counted
// 92 865 of them in one launch); one acquire fence follows the poll that succeeds.
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

// The eight epilogue warps of a CTA have finished writing a slab (their TMA stores are drained): make it visible to the team
// and count this CTA in. The CTA barrier orders every warp's stores before thread 0's fence, and a gpu-scope fence is
// cumulative over what its thread has observed (the grid-barrier idiom) -- one MEMBAR per CTA instead of 256.
__device__ __forceinline__ void team_signal(unsigned* flag, int tid) {
  epi_bar();
  if (tid == 0) {
    fence_proxy_async_global();
    fence_acq_rel_gpu();
    red_relaxed_gpu(flag, 1u);
  }
}

// Job j of the chain for a team of `team` pairs: the pair's n_pair output columns are walked in nsubt sub-tiles of w columns
// (one M = 256, N = w accumulator each; b_rows = w / 2 weight rows per CTA and 64-channel chunk).
constexpr float CH_FP8_GAIN = 8.0f;                  // extra scale of an e4m3 A operand: centres N(0, 1)-like rows in the e4m3 range
constexpr float CH_FP8_HGAIN = 16.0f;                // static scale of the e4m3 hidden activation (GELU output; saturates at 28): f5.cu folds 1 / it into gate / bias of ff2
struct JobShape { int K, N, n_pair, w, nsubt, b_rows, chunks, kstep, fp8; };
__device__ __host__ __forceinline__ JobShape job_shape(int j, int D, int FF, int team, int fp8 = 0) {
  JobShape s;
  s.K = j == 2 ? FF : D;
  s.fp8 = (fp8 && (j == 1 || j == 3)) || (fp8 >= 2 && j == 2) ? 1 : 0;      // level 1: ff1, q|k|v; level 2: ff2 as well
  s.kstep = s.fp8 ? 128 : BK;                       // operand elements per 128-byte swizzle row = per pipeline chunk
  s.chunks = s.K / s.kstep;
  s.N = j == 0 ? D : j == 1 ? FF : j == 2 ? D : 3 * D;
  s.n_pair = s.N / team;
  s.w = s.n_pair <= 256 ? s.n_pair : (s.n_pair % 256 == 0 ? 256 : 192);
  s.nsubt = s.n_pair / s.w;
  s.b_rows = s.w / 2;
  return s;
}

// First output column of sub-tile `st` of pair `slice`. Jobs 0-2: the pair's columns are contiguous (the LayerNorm partials are
// grouped per slice). Job 3 (q|k|v): sub-tiles are dealt round-robin over the team and walked from the last column down, so that
// every pair meets its V columns FIRST: their transposing epilogue (2-byte stores, ~2x the time of a q / k tile) then runs under
// the next sub-tile's main loop, and the exposed last epilogue of the launch is a q / k one (chain timeline, profiles/r02).
__device__ __forceinline__ int subtile_col0(int j, int st, int slice, int team, const JobShape& js) {
  return j == 3 ? ((js.nsubt - 1 - st) * team + slice) * js.w : slice * js.n_pair + st * js.w;
}

// This CTA's LayerNorm partials (left in shared memory by the row-layout epilogue: st[group][parity][row] = (sum, sum of
// squares), one group per 128 output columns of the CTA's slice) -> the row block's global slots. Global slot = the group's
// index among the 8 groups of a row (D = 1024): the same 8 numbers whatever the team size, reduced in the same order by every
// reader. The caller's team_signal() publishes them together with the tiles the same epilogue wrote.
__device__ __forceinline__ void publish_stats(const ChainArgs& c, int tsz, int rb, int rank, int slice, int kind, const float* st, int tid) {
  epi_bar();
  float* stats = c.stats + ((size_t)(rb * 2 + kind) * CH_ROWS + rank * 128) * (CH_MAX_TEAM * 2);
  const int groups = CH_MAX_TEAM / tsz;
  if (tid < 128) {
    for (int gi = 0; gi < groups; ++gi) {
      const float2 p0 = *reinterpret_cast<const float2*>(st + ((size_t)(gi * 2 + 0) * 128 + tid) * 2);
      const float2 p1 = *reinterpret_cast<const float2*>(st + ((size_t)(gi * 2 + 1) * 128 + tid) * 2);
      *reinterpret_cast<float2*>(stats + ((size_t)tid * CH_MAX_TEAM + slice * groups + gi) * 2) =
          make_float2(__fadd_rn(p0.x, p1.x), __fadd_rn(p0.y, p1.y));
    }
  }
}
// (mean, 1 / std) of row `r` (0 .. 127 of this CTA) from the 8 published slots; the caller has waited for the team's counter
__device__ __forceinline__ float2 row_mean_rstd(const ChainArgs& c, int rb, int rank, int kind, int r) {
  const float4* p = reinterpret_cast<const float4*>(c.stats + (((size_t)(rb * 2 + kind) * CH_ROWS + rank * 128 + r) * CH_MAX_TEAM) * 2);
  const float4 a0 = __ldcg(p), a1 = __ldcg(p + 1), a2 = __ldcg(p + 2), a3 = __ldcg(p + 3);      // (s0, q0, s1, q1) ...
  const float s = __fadd_rn(__fadd_rn(__fadd_rn(a0.x, a0.z), __fadd_rn(a1.x, a1.z)), __fadd_rn(__fadd_rn(a2.x, a2.z), __fadd_rn(a3.x, a3.z)));
  const float q = __fadd_rn(__fadd_rn(__fadd_rn(a0.y, a0.w), __fadd_rn(a1.y, a1.w)), __fadd_rn(__fadd_rn(a2.y, a2.w), __fadd_rn(a3.y, a3.w)));
  const float inv_d = 1.0f / (float)c.D;
  const float mean = __fmul_rn(s, inv_d);
  const float rstd = rsqrtf(__fadd_rn(fmaxf(__fsub_rn(__fmul_rn(q, inv_d), __fmul_rn(mean, mean)), 0.f), 1e-6f));
  return make_float2(mean, rstd);
}

// LayerNorm statistics + modulation of this CTA's [128 rows] x [D / team columns] slab of x (all 256 epilogue threads).
//   rsum / rsq : this lane's row sums over the warp's column blocks (epilogue_rows_tma<.., STATS>)
// Measured alternatives that were SLOWER on B200 (profiles/r02/chain_timeline.md): fetching the slab before the statistics
// arrive with 16 row loads per lane in flight (+3.6 ms per utterance), and walking whole rows instead of 128-column groups
// (+6 ms): more loads in flight per thread only lengthen the loaded L2 latency here.
__device__ __forceinline__ void ln_phase(const ChainArgs& c, int tsz, int rb, int rank, int slice, int kind,
                                         const float* scale, const float* shift, __nv_bfloat16* dst, unsigned* flag_stat,
                                         unsigned* flag_ready, float* st, int warp, int lane) {
  const int tid = warp * 32 + lane;
  publish_stats(c, tsz, rb, rank, slice, kind, st, tid);
  float* stats = c.stats + ((size_t)(rb * 2 + kind) * CH_ROWS + rank * 128) * (CH_MAX_TEAM * 2);
  epi_bar();
  if (tid == 0) {
    fence_acq_rel_gpu();
    red_relaxed_gpu(flag_stat, 1u);
    wait_counter(flag_stat, 2u * (unsigned)tsz);
  }
  epi_tma_drain(lane);                                         // x tiles of this warp are in global memory (under the exchange)
  epi_bar();
  if (tid == 0) stamp(c, 8 + 8 * (kind * 2) + 6);
  // pass 2: warp w normalises rows [16w, 16w + 16) of the slab, 128 columns at a time (lane l: columns 4l .. 4l + 3)
  const int n_slab = c.D / tsz;
  const float inv_d = 1.0f / (float)c.D;
  const long grow0 = (long)rb * CH_ROWS + rank * 128 + warp * 16;
  float2 pr[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {     // lane = (row k*4 + lane/8, slot lane%8)
    pr[k] = __ldcg(reinterpret_cast<const float2*>(stats + ((size_t)(warp * 16 + k * 4 + (lane >> 3)) * CH_MAX_TEAM + (lane & 7)) * 2));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      pr[k].x += __shfl_xor_sync(0xffffffffu, pr[k].x, o);
      pr[k].y += __shfl_xor_sync(0xffffffffu, pr[k].y, o);
    }
    // explicit rounding steps here and below: a row's arithmetic must not depend on the slot of the unrolled loop it falls in
    // (the compiler contracted them differently per slot: two copies of one utterance in a batch differed in the last bit)
    const float mean = __fmul_rn(pr[k].x, inv_d);
    pr[k].x = mean;                                            // -> (mean, rstd) of row k*4 + lane/8
    pr[k].y = rsqrtf(__fadd_rn(fmaxf(__fsub_rn(__fmul_rn(pr[k].y, inv_d), __fmul_rn(mean, mean)), 0.f), 1e-6f));
  }
#pragma unroll 1
  for (int cg = 0; cg < n_slab; cg += 128) {
    const int col = slice * n_slab + cg + lane * 4;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + col));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + col));
    const float4 sc1 = make_float4(__fadd_rn(1.0f, sc.x), __fadd_rn(1.0f, sc.y), __fadd_rn(1.0f, sc.z), __fadd_rn(1.0f, sc.w));
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      float4 xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long g = grow0 + half * 8 + j;
        xv[j] = g < c.R ? __ldcg(reinterpret_cast<const float4*>(c.x + g * c.D + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int jj = half * 8 + j;
        const float2 ms = half == 0 ? (j < 4 ? pr[0] : pr[1]) : (j < 4 ? pr[2] : pr[3]);
        const float mean = __shfl_sync(0xffffffffu, ms.x, (jj & 3) * 8), rstd = __shfl_sync(0xffffffffu, ms.y, (jj & 3) * 8);
        const long g = grow0 + jj;
        if (g >= c.R) continue;                                 // warp-uniform
        const float4 v = xv[j];
        *reinterpret_cast<uint2*>(dst + g * c.D + col) =
            pack16x4(__fmaf_rn(__fmul_rn(__fsub_rn(v.x, mean), rstd), sc1.x, sh.x), __fmaf_rn(__fmul_rn(__fsub_rn(v.y, mean), rstd), sc1.y, sh.y),
                     __fmaf_rn(__fmul_rn(__fsub_rn(v.z, mean), rstd), sc1.z, sh.z), __fmaf_rn(__fmul_rn(__fsub_rn(v.w, mean), rstd), sc1.w, sh.w),
                     c.f16);
      }
    }
  }
  if (tid == 0) stamp(c, 40 + kind);
  team_signal(flag_ready, tid);
}

// The two phases of a launch as seen by CTA pair `pair`: phase 0 = whole rounds of row blocks at team size c.team (every pair
// busy), phase 1 = the remainder, one team of c.team1 >= c.team pairs per block. A launch whose row blocks are one more than a
// multiple of the resident pairs used to cost a whole extra round (configs[3]: 593 = 8 x 74 + 1 blocks on one GPU, 297 / 149 / 75
// per rank on 2 / 4 / 8); now the odd block is shared by 8 pairs and costs a sixth of a round.
struct Phase { int tsz, slice, rb0, rb1, step; };
__device__ __forceinline__ bool get_phase(const ChainArgs& c, int pair, int ph, Phase& p) {
  if (ph == 0) {
    p.tsz = c.team;
    const int tm = pair / c.team;
    p.slice = pair - tm * c.team;
    p.rb0 = tm; p.rb1 = c.nrb0; p.step = c.teams;
    return tm < c.teams && tm < c.nrb0;
  }
  p.tsz = c.team1;
  const int tm = pair / c.team1;
  p.slice = pair - tm * c.team1;
  p.rb0 = c.nrb0 + tm; p.rb1 = p.rb0 + 1; p.step = 1;
  return tm < c.rem;
}

template <int FP8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS3, 1)
dit_chain_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
                 const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mA3,
                 const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1,
                 const __grid_constant__ CUtensorMap mB2, const __grid_constant__ CUtensorMap mB3,
                 const __grid_constant__ CUtensorMap mC0, const __grid_constant__ CUtensorMap mC1,      // the weight maps of phase 1
                 const __grid_constant__ CUtensorMap mC2, const __grid_constant__ CUtensorMap mC3,      // (box rows depend on the team size)
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
  uint64_t* acc_full = b_empty + CH_B_STAGES;            // [2] one per TMEM half
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint64_t* res_bar = acc_empty + 2;                     // [8] one per epilogue warp (residual tiles)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int njobs = c.has_qkv ? 4 : 3;

  if (warp == WARP_TMA && lane == 0) {
    prefetch_tmap(&mA0); prefetch_tmap(&mA1); prefetch_tmap(&mA2); prefetch_tmap(&mA3);
    prefetch_tmap(&mB0); prefetch_tmap(&mB1); prefetch_tmap(&mB2); prefetch_tmap(&mB3);
    if (c.rem > 0) { prefetch_tmap(&mC0); prefetch_tmap(&mC1); prefetch_tmap(&mC2); prefetch_tmap(&mC3); }
    prefetch_tmap(&mX); prefetch_tmap(&mFFo); prefetch_tmap(&mQKo);
    for (int s = 0; s < CH_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < CH_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(&acc_full[j], 1); mbar_init(&acc_empty[j], 16); }
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
      // ===== A producer (both CTAs): this CTA's 128 rows of each job's A operand, once per sub-tile; waits for the team's
      // hand-off before the first sub-tile of a job =====
      pdl_wait();
      int sa = 0; uint32_t pa = 0;
      for (int ph = 0; ph < 2; ++ph) {
      Phase P;
      if (!get_phase(c, pair, ph, P)) continue;
      const unsigned team_ctas = 2u * (unsigned)P.tsz;
      for (int rb = P.rb0; rb < P.rb1; rb += P.step) {
        const int row0 = rb * CH_ROWS + (int)rank * 128;
        const unsigned* flags = c.flags + (size_t)rb * CH_NFLAGS;
        for (int j = 0; j < njobs; ++j) {
          const JobShape js = job_shape(j, c.D, c.FF, P.tsz, FP8);
          const CUtensorMap* map = j == 0 ? &mA0 : j == 1 ? &mA1 : j == 2 ? &mA2 : &mA3;
          stamp(c, 8 + 8 * j + 0);
          if (j > 0) {
            wait_counter(flags + (j == 1 ? F_N16 : j == 2 ? F_FF16 : F_N16B), team_ctas);
            fence_proxy_async_global();
          }
          stamp(c, 8 + 8 * j + 1);
          for (int st = 0; st < js.nsubt; ++st)
            for (int ck = 0; ck < js.chunks; ++ck) {
              mbar_wait(&a_empty[sa], pa ^ 1);
              if (leader) mbar_expect_tx(&a_full[sa], 2u * CH_A_BYTES);
              tma_load_3d_2sm(smem_a + sa * CH_A_BYTES, map, &a_full[sa], ck * js.kstep, row0, 0);
              if (++sa == CH_A_STAGES) { sa = 0; pa ^= 1; }
            }
        }
      }
      }
    }
  } else if (warp == WARP_TMA_B) {
    if (lane == 0) {
      // ===== B producer (both CTAs): this CTA's half of each sub-tile's weight rows. No pdl_wait: weights are constants =====
      int sb = 0; uint32_t pb = 0;
      for (int ph = 0; ph < 2; ++ph) {
      Phase P;
      if (!get_phase(c, pair, ph, P)) continue;
      const int slice = P.slice;
      for (int rb = P.rb0; rb < P.rb1; rb += P.step) {
        for (int j = 0; j < njobs; ++j) {
          const JobShape js = job_shape(j, c.D, c.FF, P.tsz, FP8);
          const CUtensorMap* map = ph == 0 ? (j == 0 ? &mB0 : j == 1 ? &mB1 : j == 2 ? &mB2 : &mB3)
                                           : (j == 0 ? &mC0 : j == 1 ? &mC1 : j == 2 ? &mC2 : &mC3);
          const uint32_t tx = 2u * (uint32_t)(js.b_rows * 128);
          for (int st = 0; st < js.nsubt; ++st) {
            const int n_row0 = subtile_col0(j, st, slice, P.tsz, js) + (int)rank * js.b_rows;
            for (int ck = 0; ck < js.chunks; ++ck) {
              mbar_wait(&b_empty[sb], pb ^ 1);
              if (leader) mbar_expect_tx(&b_full[sb], tx);
              tma_load_3d_2sm(smem_b + sb * CH_B_BYTES, map, &b_full[sb], ck * js.kstep, n_row0, 0);
              if (++sb == CH_B_STAGES) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
      }
    }
  } else if (warp == WARP_MMA) {
    if (leader) {
      // ===== MMA issuer (leader CTA only): M = 256 across the pair; sub-tile t accumulates in TMEM half t & 1 =====
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(smem_a)), b_lo0 = desc_lo_sw128(smem_u32(smem_b));
      const uint32_t a_stage_lo = (uint32_t)CH_A_BYTES >> 4, b_stage_lo = (uint32_t)CH_B_BYTES >> 4;
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      uint32_t t = 0;
      for (int ph = 0; ph < 2; ++ph) {
      Phase P;
      if (!get_phase(c, pair, ph, P)) continue;
      for (int rb = P.rb0; rb < P.rb1; rb += P.step) {
        for (int j = 0; j < njobs; ++j) {
          const JobShape js = job_shape(j, c.D, c.FF, P.tsz, FP8);
          const uint32_t idesc = idesc_f16kind(256, js.w, js.fp8 ? 1 : c.f16);      // e4m3 shares format code 0 with fp16
          const int chunks = js.chunks;
          for (int st = 0; st < js.nsubt; ++st, ++t) {
            const uint32_t buf = t & 1u;
            mbar_wait(&acc_empty[buf], ((t >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d0 = tmem_base + buf * 256u;
            for (int ck = 0; ck < chunks; ++ck) {
              mbar_wait(&a_full[sa], pa);
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              if (ck == 0 && st == 0 && lane == 0) stamp(c, 8 + 8 * j + 2);
              const uint32_t a_lo = a_lo0 + (uint32_t)sa * a_stage_lo;
              const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_stage_lo;
              const uint32_t accum = ck > 0 ? 1u : 0u;
              if (elect_one()) {
                if (FP8 && js.fp8) {
                  umma2_f8_lohi(d0, a_lo + 0, b_lo + 0, idesc, accum);
                  umma2_f8_lohi(d0, a_lo + 2, b_lo + 2, idesc, 1u);
                  umma2_f8_lohi(d0, a_lo + 4, b_lo + 4, idesc, 1u);
                  umma2_f8_lohi(d0, a_lo + 6, b_lo + 6, idesc, 1u);
                } else {
                  umma2_bf16_lohi(d0, a_lo + 0, b_lo + 0, idesc, accum);
                  umma2_bf16_lohi(d0, a_lo + 2, b_lo + 2, idesc, 1u);
                  umma2_bf16_lohi(d0, a_lo + 4, b_lo + 4, idesc, 1u);
                  umma2_bf16_lohi(d0, a_lo + 6, b_lo + 6, idesc, 1u);
                }
                umma2_commit_mc(&b_empty[sb]);
                umma2_commit_mc(&a_empty[sa]);
                if (ck == chunks - 1) { umma2_commit_mc(&acc_full[buf]); if (st == js.nsubt - 1) stamp(c, 8 + 8 * j + 3); }
              }
              __syncwarp();
              if (++sa == CH_A_STAGES) { sa = 0; pa ^= 1; }
              if (++sb == CH_B_STAGES) { sb = 0; pb ^= 1; }
            }
          }
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
    float* s_gate = smem_vec + 256;
    float* s_mul = smem_vec + 512;
    // bias (or v) / gate (or u) / 1 + scale of the sub-tile's columns -> shared memory (broadcast reads in the epilogue)
    auto stage_vec = [&](const float* bias, const float* gate, const float* scale, const float* wscale, int n0, int n) {
      epi_bar();                                               // the previous sub-tile's readers are done
      if (tid < n) {
        s_bias[tid] = __ldg(bias + n0 + tid);
        if (gate) s_gate[tid] = __ldg(gate + n0 + tid);
        if (scale) s_mul[tid] = __fadd_rn(1.0f, __ldg(scale + n0 + tid));
        if (wscale) s_mul[tid] = __ldg(wscale + n0 + tid);     // (a job has one or the other)
      }
      epi_bar();
    };
    TcArgs a;
    a.taps = 1; a.dil = 1; a.center = 0; a.groups = 1; a.M = c.R;
    a.o_bstride = 0; a.o_shift = 0; a.accumulate = 0; a.scale = 1.0f; a.out2 = nullptr; a.f16 = c.f16;
    a.rope_cs = nullptr; a.rope_cols = 0; a.rope_rows = 1; a.vt_out = nullptr; a.vt_col0 = 0; a.vt_ld = 0; a.vt_heads = 0;
    a.out = nullptr; a.ldo = 0; a.o_limit = 0; a.out_bf16 = 0; a.bias = nullptr; a.gate = nullptr; a.res = nullptr;
    a.rowinfo = c.rowinfo;
    uint32_t t = 0;
    const uint64_t x_policy = c.l2hint ? l2_policy_evict_last() : 0ull;
    for (int ph = 0; ph < 2; ++ph) {
    Phase P;
    if (!get_phase(c, pair, ph, P)) continue;
    const int slice = P.slice;
    const unsigned team_ctas = 2u * (unsigned)P.tsz;
    for (int rb = P.rb0; rb < P.rb1; rb += P.step) {
      unsigned* flags = c.flags + (size_t)rb * CH_NFLAGS;
      const int row0 = rb * CH_ROWS + (int)rank * 128 + q * 32;
      const int myrow = row0 + lane;                           // the row this lane owns in every epilogue of the row block
      // LayerNorm folded into the GEMMs: job 0 / 2 write c * x * (1 + scale) next to x, job 1 / 3 undo c and apply mean / rstd.
      // c = the row's 1 / std at the PREVIOUS LayerNorm (a scale only: any c > 0 gives the same result up to rounding; a stale
      // estimate keeps the 16-bit operand in the range the normalised rows would have)
      float c_row = myrow < c.R ? __ldcg(c.rowscale + myrow) : 1.0f;
      float rho = 1.0f, rmu = 0.0f;
      for (int j = 0; j < njobs; ++j) {
        const JobShape js = job_shape(j, c.D, c.FF, P.tsz, FP8);
        a.Cin = js.K; a.N = js.N; a.BN = js.w; a.kchunks = js.chunks;
        const float* wscale = !js.fp8 || j == 2 ? nullptr : j == 1 ? c.sw_ff1 : c.sw_qkv;      // (ff2: folded into its gate / bias vectors)
        const bool next_fp8 = FP8 && (j == 0 || j == 2);      // the operand this job emits feeds an e4m3 GEMM
        const bool fold = j == 0 || (j == 2 && c.has_qkv);      // this job's epilogue emits the next GEMM's A operand
        const float* bias = j == 0 ? c.b_out : j == 1 ? c.v_ff1 : j == 2 ? c.b_ff2 : c.v_qkv;
        const float* gate = j == 0 ? c.gate_msa : j == 1 ? c.u_ff1 : j == 2 ? c.gate_mlp : c.u_qkv;
        const float* scale = !fold ? nullptr : j == 0 ? c.scale_mlp : c.scale_nxt;
        if (j == 3) {
          a.rope_cs = c.rope_cs; a.rope_cols = 2 * c.D; a.rope_rows = c.rope_rows;
          a.vt_out = c.vt_out; a.vt_col0 = 2 * c.D; a.vt_ld = c.vt_ld; a.vt_heads = c.vt_heads;
        } else {
          a.rope_cs = nullptr; a.rope_cols = 0; a.rope_rows = 1; a.vt_out = nullptr; a.vt_col0 = 0;
        }
        if (j == 1 || j == 3) {
          // the statistics of the LayerNorm in front of this GEMM: published by the whole team before it signalled the A tiles
          if (tid == 0) wait_counter(flags + (j == 1 ? F_N16 : F_N16B), team_ctas);
          epi_bar();
          const float2 ms = row_mean_rstd(c, rb, (int)rank, j == 1 ? 0 : 1, q * 32 + lane);
          rho = __fdiv_rn(ms.y, js.fp8 ? c_row * CH_FP8_GAIN : c_row);
          rmu = __fmul_rn(ms.y, ms.x);
          c_row = ms.y;                                         // the scale of the NEXT emitted operand
          if (j == 3 && slice == 0 && e == 0 && myrow < c.R) c.rowscale[myrow] = ms.y;     // for the next block's launch
        }
        EpiEmit emit;
        if (fold) {
          if (next_fp8) emit.row8 = myrow < c.R ? reinterpret_cast<uint8_t*>(j == 0 ? c.n16 : c.n16b) + (size_t)myrow * c.D : nullptr;
          else emit.row = myrow < c.R ? reinterpret_cast<uint16_t*>(j == 0 ? c.n16 : c.n16b) + (size_t)myrow * c.D : nullptr;
          emit.s_mul = s_mul; emit.c = next_fp8 ? c_row * CH_FP8_GAIN : c_row;
        } else if (wscale) {
          emit.s_mul = s_mul;                                    // AFFINE epilogue: the e4m3 weights' per-column scales
        }
        emit.c = fold ? emit.c : 0.0f;                           // (non-fold jobs: 0 = no e4m3 hidden activation)
        if (FP8 >= 2 && j == 1) {                                // ff1 writes the hidden activation as e4m3 for an e4m3 ff2
          emit.c = CH_FP8_HGAIN;
          emit.row8 = myrow < c.R ? reinterpret_cast<uint8_t*>(c.ff16) + (size_t)myrow * c.FF : nullptr;
        }
        float rsum = 0.f, rsq = 0.f;
        for (int st = 0; st < js.nsubt; ++st, ++t) {
          const uint32_t buf = t & 1u;
          const int n0 = subtile_col0(j, st, slice, P.tsz, js);
          const uint32_t taddr = tmem_base + buf * 256u + lane_sel;
          stage_vec(bias, gate, scale, wscale, n0, js.w);
          if (j == 0 || j == 2) epi_tma_fetch_res(&mX, et, n0 + e * 32, row0, lane);
          mbar_wait(&acc_full[buf], (t >> 1) & 1u);
          tc_fence_after();
          if (tid == 0 && st == 0) stamp(c, 8 + 8 * j + 4);
          if (j == 0 || j == 2)
            epilogue_rows_tma<TK_RES_F32, ACT_NONE, true, false, (FP8 != 0)>(a, &mX, &mX, et, taddr, row0, n0, e * 32, 64, lane, s_bias, s_gate, rsum, rsq, x_policy,
                                                          smem_stat, slice * js.n_pair, q * 32 + lane, 1.0f, 0.0f, emit);
          else if (j == 1)
            epilogue_rows_tma<TK_ACT16, ACT_GELU_TANH, false, true, (FP8 != 0)>(a, &mFFo, nullptr, et, taddr, row0, n0, e * 32, 64, lane, s_bias, s_gate, rsum, rsq,
                                                                    0ull, nullptr, 0, 0, rho, rmu, emit);
          else
            epilogue_rows_tma<TK_ROPE16, ACT_NONE, false, true, (FP8 != 0)>(a, &mQKo, nullptr, et, taddr, row0, n0, e * 32, 64, lane, s_bias, s_gate, rsum, rsq,
                                                                0ull, nullptr, 0, 0, rho, rmu, emit);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);
        }
        if (tid == 0) stamp(c, 8 + 8 * j + 5);
        if (fold) {
          // x tiles (TMA stores) and the emitted 16-bit rows (plain stores) of this CTA are complete: publish the LayerNorm
          // partials and hand everything to the team with ONE counter (the next GEMM's A producer and its epilogue wait on it)
          // (the x tiles still in flight through the TMA unit are nobody else's business: only this warp reads them back, two
          // jobs later; they are drained after the signal, off the team's critical path)
          publish_stats(c, P.tsz, rb, (int)rank, slice, j == 0 ? 0 : 1, smem_stat, tid);
          team_signal(flags + (j == 0 ? F_N16 : F_N16B), tid);
          epi_tma_drain(lane);
        } else if (j == 2) {
          // last block: the final LayerNorm feeds proj_out (a separate GEMM) -> normalise in a second pass as before
          ln_phase(c, P.tsz, rb, (int)rank, slice, 1, c.scale_nxt, c.shift_nxt, c.n16b, flags + F_STAT2, flags + F_N16B, smem_stat, warp, lane);
        } else {
          epi_tma_drain(lane);
          if (j == 1) team_signal(flags + F_FF16, tid);
        }
        if (tid == 0) stamp(c, 8 + 8 * j + 7);
      }
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
  static PerDeviceOnce once;
  if (once.first()) {
    B2_CUDA(cudaFuncSetAttribute(dit_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
    B2_CUDA(cudaFuncSetAttribute(dit_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
    B2_CUDA(cudaFuncSetAttribute(dit_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
  }
  int dev = 0, sms = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return sms / 2;
}

// pairs per row block: the split that finishes nrb row blocks soonest on `pairs` resident pairs (a pair's share of a block's
// work is 1 / team, and blocks beyond pairs / team wait for a second round)
int pick_team(int nrb, int pairs) {
  const char* ev = getenv("B200TTS_CHAIN_TEAM");             // experiment override
  if (ev != nullptr) { const int v = atoi(ev); if (v == 1 || v == 2 || v == 4 || v == 8) return v; }
  int best = 1;
  double best_cost = 1e300;
  for (int tm : {1, 2, 4, 8}) {                              // ties go to the smaller team: fewer hand-offs, more overlap
    const int teams = pairs / tm;
    if (teams < 1) continue;
    const double cost = (double)ceil_div(nrb, teams) / tm;
    if (cost < best_cost - 1e-12) { best_cost = cost; best = tm; }
  }
  return best;
}

}  // namespace

// plan: phase 0 = whole rounds of row blocks at team size t0, phase 1 = the remaining blocks with the largest team that still
// gives each its own team; the (t0, t1) pair with the lowest modelled time wins. Pure host arithmetic (tests/test_chain_plan_cpu.py).
DitChainPlan dit_chain_plan(int nrb, int pairs) {
  DitChainPlan pl;
  pl.team = 1; pl.teams = nrb < pairs ? nrb : pairs; pl.nrb0 = nrb; pl.rem = 0; pl.team1 = 1; pl.cost = 1e300;
  if (nrb <= 0 || pairs <= 0) return pl;
  // time of one round of row blocks by team size, relative (measured: a round of 74 blocks at T = 1 takes 340 us, a block
  // shared by 8 pairs 57 us; profiles/r02 chain timelines)
  static const double round_cost[4] = {1.0, 0.55, 0.30, 0.17};
  auto lg = [](int t) { return t == 1 ? 0 : t == 2 ? 1 : t == 4 ? 2 : 3; };
  auto team_for = [&](int blocks, int tmin) {                    // largest team size that gives every block its own team
    int t = CH_MAX_TEAM;
    while (t > tmin && blocks > pairs / t) t /= 2;
    return t;
  };
  for (int t0 = 1; t0 <= CH_MAX_TEAM; t0 *= 2) {
    const int teams0 = pairs / t0;
    if (teams0 < 1) break;
    const int full = nrb / teams0, rem = nrb - full * teams0;
    const int t1 = rem > 0 ? team_for(rem, t0) : t0;
    if (rem > pairs / t1) continue;                              // (only for t0 = 8: the remainder is another round)
    const double cost = full * round_cost[lg(t0)] + (rem > 0 ? round_cost[lg(t1)] : 0.0);
    if (cost < pl.cost - 1e-9) {
      pl.cost = cost;
      if (full == 0) { pl.team = t1; pl.teams = rem; pl.nrb0 = nrb; pl.rem = 0; pl.team1 = 1; }
      else { pl.team = t0; pl.teams = teams0; pl.nrb0 = full * teams0; pl.rem = rem; pl.team1 = t1; }
    }
  }
  return pl;
}

size_t dit_chain_stats_floats(int R) { return (size_t)ceil_div(R, CH_ROWS) * 2 * CH_ROWS * CH_MAX_TEAM * 2; }
size_t dit_chain_flag_words(int R) { return (size_t)ceil_div(R, CH_ROWS) * CH_NFLAGS; }

bool dit_chain_supported(int D, int FF, int H) {
  // the job table (sub-tiles of 128 / 192 / 256 columns, two 256-column TMEM halves) is laid out for the F5 DiT
  if (D != 1024 || FF != 2048 || H * 64 != D) return false;
  return resident_pairs() >= CH_MAX_TEAM;
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
  const int pairs = resident_pairs();
  if (getenv("B200TTS_CHAIN_TEAM") != nullptr) {                  // experiment override: one phase at the forced team size
    c.nrb0 = c.nrb; c.rem = 0; c.team1 = 1;
    c.team = pick_team(c.nrb, pairs);
    const int max_teams = pairs / c.team;
    c.teams = c.nrb < max_teams ? c.nrb : max_teams;
  } else {
    const DitChainPlan pl = dit_chain_plan(c.nrb, pairs);
    c.team = pl.team; c.teams = pl.teams; c.nrb0 = pl.nrb0; c.rem = pl.rem; c.team1 = pl.team1;
  }
  c.x = d.x; c.n16 = d.n16; c.ff16 = d.ff16; c.n16b = d.n16b;
  c.b_out = d.b_out; c.gate_msa = d.gate_msa; c.shift_mlp = d.shift_mlp; c.scale_mlp = d.scale_mlp; c.b_ff1 = d.b_ff1; c.b_ff2 = d.b_ff2;
  c.gate_mlp = d.gate_mlp; c.shift_nxt = d.shift_nxt; c.scale_nxt = d.scale_nxt; c.b_qkv = d.b_qkv;
  c.qk16 = d.qk16; c.rope_cs = d.rope_cs; c.rope_rows = d.rope_rows > 0 ? d.rope_rows : 1; c.rowinfo = d.rowinfo; c.vt_out = d.vt_out; c.vt_ld = d.vt_ld; c.vt_heads = d.vt_heads;
  B2_CHECK(d.u_ff1 && d.v_ff1 && d.rowscale && (!d.has_qkv || (d.u_qkv && d.v_qkv)), "dit_chain: folded-LayerNorm vectors");
  B2_CHECK(!d.fp8 || (d.w8_ff1 && d.sw_ff1 && (!d.has_qkv || (d.w8_qkv && d.sw_qkv))), "dit_chain: e4m3 weights / scales");
  B2_CHECK(d.fp8 < 2 || d.w8_ff2, "dit_chain: e4m3 ff2 weights");
  c.fp8 = d.fp8; c.sw_ff1 = d.sw_ff1; c.sw_qkv = d.sw_qkv;
  c.u_ff1 = d.u_ff1; c.v_ff1 = d.v_ff1; c.u_qkv = d.u_qkv; c.v_qkv = d.v_qkv; c.rowscale = d.rowscale;
  c.stats = d.stats; c.flags = d.flags; c.trace = d.trace;
  { const char* v = getenv("B200TTS_CHAIN_L2HINT"); c.l2hint = v != nullptr && atoi(v) != 0; }
  CUtensorMap mA[4], mB[4];
  const void* a_ptr[4] = {d.att16, d.n16, d.ff16, d.n16b};
  const int a_k[4] = {d.D, d.D, d.FF, d.D};
  for (int j = 0; j < 4; ++j) {
    if ((d.fp8 && (j == 1 || j == 3)) || (d.fp8 >= 2 && j == 2))     // e4m3 rows of K bytes in the 16-bit rows' buffer, densely packed
      tc_encode_map_u8(&mA[j], a_ptr[j], (uint64_t)a_k[j], (uint64_t)d.R, 1, (uint64_t)a_k[j], (uint64_t)d.R * a_k[j], 128);
    else
      tc_encode_map(&mA[j], a_ptr[j], (uint64_t)a_k[j], (uint64_t)d.R, 1, (uint64_t)a_k[j], (uint64_t)d.R * a_k[j], 128);
  }
  const TcWeight* ws[4] = {d.w_out, d.w_ff1, d.w_ff2, wq};
  CUtensorMap mC[4];
  for (int ph = 0; ph < 2; ++ph)
    for (int j = 0; j < 4; ++j) {
      CUtensorMap* mp = ph == 0 ? &mB[j] : &mC[j];
      const JobShape js = job_shape(j == 3 && !d.has_qkv ? 0 : j, d.D, d.FF, ph == 0 ? c.team : c.team1, d.fp8);
      if ((d.fp8 && (j == 1 || (j == 3 && d.has_qkv))) || (d.fp8 >= 2 && j == 2)) {
        const void* w8 = j == 1 ? d.w8_ff1 : j == 2 ? d.w8_ff2 : d.w8_qkv;
        const uint64_t n_rows = j == 1 ? (uint64_t)d.FF : j == 2 ? (uint64_t)d.D : (uint64_t)3 * d.D;
        const uint64_t kk = j == 2 ? (uint64_t)d.FF : (uint64_t)d.D;
        tc_encode_map_u8(mp, w8, kk, n_rows, 1, kk, n_rows * kk, (uint32_t)js.b_rows);
      } else {
        tc_encode_map(mp, ws[j]->w.p, (uint64_t)ws[j]->Cin, (uint64_t)ws[j]->N, 1, (uint64_t)ws[j]->ldc, (uint64_t)ws[j]->N * ws[j]->ldc,
                      (uint32_t)js.b_rows);
      }
    }
  // epilogue tiles: x (fp32, residual in / result out), ff16 and q|k (16 bit, result out), 32 rows x 32 columns each
  CUtensorMap mX, mFFo, mQKo;
  tc_encode_map2d(&mX, d.x, 4, (uint64_t)d.D, (uint64_t)d.R, (uint64_t)d.D, 32, 32);
  tc_encode_map2d(&mFFo, d.ff16, 2, (uint64_t)d.FF, (uint64_t)d.R, (uint64_t)d.FF, 32, 32);
  if (d.has_qkv) tc_encode_map2d(&mQKo, d.qk16, 2, (uint64_t)2 * d.D, (uint64_t)d.R, (uint64_t)2 * d.D, 32, 32);
  else mQKo = mFFo;
  const int grid_pairs = std::max(c.teams * c.team, c.rem * c.team1);
  auto kern = d.fp8 >= 2 ? dit_chain_kernel<2> : d.fp8 == 1 ? dit_chain_kernel<1> : dit_chain_kernel<0>;
  launch_pdl(kern, dim3((unsigned)(grid_pairs * 2)), dim3(NTHREADS3), (size_t)CH_SMEM, stream, mA[0], mA[1], mA[2], mA[3],
             mB[0], mB[1], mB[2], mB[3], mC[0], mC[1], mC[2], mC[3], mX, mFFo, mQKo, c);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts

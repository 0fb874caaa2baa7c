// Non-causal, unmasked, unscaled flash attention for the F5 DiT (modules.py:449-468: softmax(q @ k) @ v in
// fp32, the 1/sqrt(d) scale already folded into Wq/Wk at export) on tcgen05 tensor cores.
//
// One CTA = one (batch, head, 128-query tile). Per 128-key block j:
//   S_j  = Q K_j^T   tcgen05.mma  M=128 N=128 K=64   (Q, K tiles: TMA boxes of the [2N][2048] q|k tensor, SWIZZLE_128B)
//   P_j  = exp2(S_j log2e - m)   256 softmax threads: thread (row r, key half g) owns 64 keys of its row in ONE pass
//          (the 64 scores are pulled from TMEM into registers, which also frees the S buffer so that S_{j+1} is computed
//          while the exponentials run), packed FFMA2 / FADD2 arithmetic, P (16 bit) back into TENSOR MEMORY
//   O   += P_j V_j   tcgen05.mma  M=128 N=64 K=128 with the A operand in tensor memory, accumulated in TMEM across blocks.
// Round 2 (profiles/r02/ncu_summary_r02a.md): P used to go through ONE shared-memory tile, so the softmax threads of block j+1
// had to wait for P_j V_j to finish before they could write (p_full -> MMA issue -> pv_done = 0.5 us of a 1.1-1.5 us block
// period; XU pipe 49 % busy). Now the probabilities stay in registers until the end of the block and only then wait for
// P_{j-1} V_{j-1} (long done); the MMA runs under the next block's exponentials and the softmax warps never stall on it.
// The two threads of a row share ONE running maximum (exchanged through shared memory, a 64-thread named barrier per warp
// pair), hence one O accumulator: TMEM = S 128 + O 64 + P 64 columns = 256 -> two CTAs per SM.
// The running maximum is lazy (FlashAttention-4 style): it only moves when a block's maximum exceeds it by more than 8
// (log2 units), and only then is O rescaled in TMEM (tcgen05.ld / multiply / tcgen05.st, warp-voted); P stays below
// 2^8, which bf16 / fp16 and the fp32 accumulators hold without loss.
// A fraction of the exponentials (template POLY of every 4 pairs) runs on the FMA pipes instead of the MUFU unit
// (Cody-Waite range reduction + degree-3 minimax polynomial, relative error 8.8e-5 < half an ulp of fp16): the kernel is
// MUFU-bound otherwise (16 ex2 per clock and SM against 8192 tensor FLOPs).
// Threads: 4 G softmax / epilogue warps (G = 2 threads per query row by default, 4 optional), then the TMA producer and the
// MMA issuer + TMEM owner. Shared memory 84 KB and 256 TMEM columns per CTA -> two CTAs per SM.
// What bounds it (profiles/r02/attention_r02.md): not one pipe. Removing the exponentials, the MMAs and the loads altogether
// (debug switches, since deleted) only shortened a launch from 20.2 to 15.9 us: ~380 instructions per warp and key block at an
// IPC of 0.5 per scheduler (stall reasons: fixed-latency `wait` 24 %, scoreboard 27 %, MIO / math throttle 10 %), i.e. issue-
// and latency-bound softmax code. 4 threads per row (twice the warps, half the work each) measured within noise of 2; half of the
// exponentials on the FMA pipes (POLY = 2, the default) is worth 5 % of the kernel in the batched benchmark. Both are switches
// (B200TTS_ATTN_G, B200TTS_ATTN_POLY).
#include "attention_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "rowgemm_tc.cuh"
#include "tc_ptx.cuh"

namespace b200tts {

namespace {

using namespace tc;

constexpr int BQ = 128, BKEY = 128, HD = 64;
// G softmax threads per query row (2 or 4): thread (row r, key group g) owns 128 / G keys of every block. 4 G softmax warps,
// then the TMA producer and the MMA issuer (highest warp ids: the SMSP arbiter favours them over the softmax warps)
constexpr int nthreads(int G) { return (4 * G + 2) * 32; }
constexpr int Q_BYTES = BQ * HD * 2;            // 16 KB
constexpr int K_BYTES = BKEY * HD * 2;          // 16 KB
constexpr int V_BYTES = HD * BKEY * 2;          // 16 KB = two [64 d][64 keys] chunks
constexpr int XCH_BYTES = 2 * 4 * BQ * 4;       // [block parity][key group][row] floats: the row maximum / row sum exchange
constexpr int KV_STAGES = 2;
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + XCH_BYTES + 128;      // 84 KB
constexpr float RESCALE_TAU = 8.0f;             // log2 units
constexpr int ATTN_POLY_DEFAULT = 2;      // configs[3]: 3569 ms per step against 3617 (POLY 1) / 3609 (POLY 0), profiles/r02/probe_zl
constexpr int ATTN_G_DEFAULT = 2;
// tensor-memory columns of a CTA
constexpr uint32_t TM_S = 0, TM_O = 128, TM_P = 192, TM_COLS = 256;

struct AttnArgs {
  int N, H;
  __nv_bfloat16* out;
  int ldo;
  const int* seq_off; // ragged batches: first row of every sequence in the concatenated [rows][2*H*64] tensor (null: b * N)
  const int* seq_len; // ragged batches: tokens of every sequence (null: N)
  unsigned long long* trace;   // debug: [CTA][64] %globaltimer stamps (B200TTS_ATTN_TRACE=<file>, tools/attn_trace.py)
};

__device__ __forceinline__ void astamp(const AttnArgs& a, int slot) {
  if (a.trace != nullptr && slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    a.trace[cta * 64 + slot] = t;
  }
}

template <bool F16>
__device__ __forceinline__ uint32_t pk16(float x, float y) {
  if (F16) { __half2 h = __floats2half2_rn(x, y); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 p = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
// packed fp32 pairs (FFMA2 / FADD2)
__device__ __forceinline__ unsigned long long pk2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void up2(unsigned long long r, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long add2_rm(unsigned long long a, unsigned long long b) {      // round towards -inf
  unsigned long long d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^e for a pair of exponents e <= 8 on the FMA pipes: e = n + f with n = floor(e) (the magic-number add in round-down mode
// leaves n in the low mantissa bits), 2^f by a degree-3 minimax polynomial on [0, 1) (relative error 8.8e-5), n added to
// the exponent field by integer arithmetic. Exponents below -126 are clamped (the result is then a denormal ~ 1e-38).
__device__ __forceinline__ void ex2_fma_pair(float e0, float e1, float& p0, float& p1) {
  const float MAGIC = 12582912.f;               // 1.5 * 2^23
  e0 = fmaxf(e0, -126.f);
  e1 = fmaxf(e1, -126.f);
  const unsigned long long x2 = pk2(e0, e1);
  const unsigned long long t2 = add2_rm(x2, pk2(MAGIC, MAGIC));
  const unsigned long long n2 = add2(t2, pk2(-MAGIC, -MAGIC));                   // floor(e), exact
  const unsigned long long f2 = fma2(n2, pk2(-1.f, -1.f), x2);                   // e - floor(e) in [0, 1)
  unsigned long long q2 = fma2(f2, pk2(0.077119089663028717f, 0.077119089663028717f), pk2(0.227564394474029541f, 0.227564394474029541f));
  q2 = fma2(q2, f2, pk2(0.695146143436431885f, 0.695146143436431885f));
  q2 = fma2(q2, f2, pk2(1.f, 1.f));
  float t0, t1, q0, q1;
  up2(t2, t0, t1);
  up2(q2, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// the G warps (key groups) that share query rows [32 qd, 32 qd + 32)
template <int G>
__device__ __forceinline__ void rows_bar(int qd) {
  switch (qd) {                                   // immediate barrier ids: a register id makes ptxas reserve all 16
    case 0: asm volatile("bar.sync 1, %0;" ::"n"(32 * G) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"n"(32 * G) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"n"(32 * G) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"n"(32 * G) : "memory"); break;
  }
}
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st32(taddr, r); }
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = lanes, K = 16 elements = 8 columns of packed 16-bit pairs)
// is read from tensor memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// POLY: of every 4 score pairs, POLY are exponentiated on the FMA pipes (0 .. 3)
template <bool F16, int POLY, int G>
__global__ void __launch_bounds__(nthreads(G), 2) attn_tc_kernel(const __grid_constant__ CUtensorMap map_qk,
                                                              const __grid_constant__ CUtensorMap map_v, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KV_STAGES * K_BYTES;
  float* xch = reinterpret_cast<float*>(sV + KV_STAGES * V_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xch) + XCH_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                  // [2]  K and V stages are released separately: a K tile is dead as soon as
  uint64_t* k_empty = bars + 3;                 // [2]  S_j = Q K_j^T has run, its V tile only after O += P_j V_j, a whole
  uint64_t* v_full = bars + 5;                  // [2]  softmax later
  uint64_t* v_empty = bars + 7;                 // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;                 // softmax threads hold S in registers: the S buffer may be overwritten
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  constexpr int WARP_TMA = 4 * G, WARP_MMA = 4 * G + 1;
  constexpr int KT = BKEY / G;                    // keys per softmax thread and block
  constexpr int OC = HD / G;                      // output columns per softmax thread
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y, b = blockIdx.z;
  // ragged batch: sequence b has Nb tokens starting at row rb0 of ONE concatenated tensor (TMA batch coordinate 0); rows past
  // its end belong to the next sequence and are masked like the keys of a ragged last block
  const int Nb = a.seq_len ? __ldg(a.seq_len + b) : a.N;
  const int rb0 = a.seq_off ? __ldg(a.seq_off + b) : 0;
  const int bz = a.seq_off ? 0 : b;
  if (q0 >= Nb) return;                           // grid.x covers the longest sequence (whole CTA leaves: nothing allocated yet)
  const int nblocks = (Nb + BKEY - 1) / BKEY;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();       // SWIZZLE_128B tiles need a 1024-byte aligned base
    prefetch_tmap(&map_qk);
    prefetch_tmap(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4 * G);                   // one arrival per softmax WARP: 256 per-thread arrivals on one barrier word
    mbar_init(p_full, 4 * G);                   // serialise in the shared-memory atomic unit (~1 us per block, profiles/r02)
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc(tmem_ptr, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();                                          // after the TMEM allocation (common.cuh)
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) astamp(a, 0);
  const uint32_t tmem_S = tmem_base + TM_S, tmem_O = tmem_base + TM_O, tmem_P = tmem_base + TM_P;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_3d(sQ, &map_qk, q_full, h * HD, rb0 + q0, bz);
      const int bh = b * a.H + h;
      for (int j = 0; j < nblocks; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (uint32_t)(j / KV_STAGES) & 1u;
        mbar_wait(&k_empty[s], ph ^ 1u);
        mbar_expect_tx(&k_full[s], K_BYTES);
        tma_load_3d(sK + s * K_BYTES, &map_qk, &k_full[s], a.H * HD + h * HD, rb0 + j * BKEY, bz);
        mbar_wait(&v_empty[s], ph ^ 1u);
        mbar_expect_tx(&v_full[s], V_BYTES);
        tma_load_3d(sV + s * V_BYTES, &map_v, &v_full[s], j * BKEY, 0, bh);
        tma_load_3d(sV + s * V_BYTES + V_BYTES / 2, &map_v, &v_full[s], j * BKEY + 64, 0, bh);
      }
    }
  } else if (warp == WARP_MMA) {
    if (lane == 0) {
      const uint32_t fmt = F16 ? 0u : ((1u << 7) | (1u << 10));          // kind::f16 operand format: 1 = bf16, 0 = fp16
      const uint32_t idesc_s = (1u << 4) | fmt | ((uint32_t)(BKEY >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | fmt | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint64_t dQ = make_desc_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const int s = j % KV_STAGES;
        mbar_wait(&k_full[s], (uint32_t)(j / KV_STAGES) & 1u);
        if (j > 0) mbar_wait(s_free, (uint32_t)(j - 1) & 1u);       // S_{j-1} is in the softmax threads' registers
        tc_fence_after();
        const uint64_t dK = make_desc_sw128(smem_u32(sK + s * K_BYTES));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_S, dQ + (uint64_t)(2 * k), dK + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&k_empty[s]);
        astamp(a, 4 + 4 * j + 0);                                   // the K tile is free once this MMA has read it
      };
      mbar_wait(q_full, 0);
      astamp(a, 1);
      issue_s(0);
      for (int j = 0; j < nblocks; ++j) {
        if (j + 1 < nblocks) issue_s(j + 1);                        // runs under the exponentials of block j
        const int s = j % KV_STAGES;
        mbar_wait(&v_full[s], (uint32_t)(j / KV_STAGES) & 1u);
        mbar_wait(p_full, (uint32_t)j & 1u);
        tc_fence_after();
        const uint64_t dV0 = make_desc_sw128(smem_u32(sV + s * V_BYTES));
        const uint64_t dV1 = make_desc_sw128(smem_u32(sV + s * V_BYTES + V_BYTES / 2));
#pragma unroll
        for (int k = 0; k < 4; ++k)                                 // keys [0, 64) of the block: P columns 0..31
          umma_ts(tmem_O, tmem_P + (uint32_t)(8 * k), dV0 + (uint64_t)(2 * k), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)                                 // keys [64, 128): P columns 32..63
          umma_ts(tmem_O, tmem_P + (uint32_t)(32 + 8 * k), dV1 + (uint64_t)(2 * k), idesc_pv, 1u);
        umma_commit(pv_done);
        umma_commit(&v_empty[s]);
        astamp(a, 4 + 4 * j + 1);
      }
    }
  } else {
    // ===== softmax + epilogue: row r of the query tile, key group g (KT keys) of every block =====
    const int qd = warp & 3, g = warp >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_Sg = tmem_S + lane_addr + (uint32_t)(g * KT);            // this thread's KT scores
    const uint32_t tmem_Og = tmem_O + lane_addr + (uint32_t)(g * OC);            // its OC of the 64 output columns
    const uint32_t tmem_Pg = tmem_P + lane_addr + (uint32_t)(g * KT / 2);        // its KT probabilities = KT / 2 packed columns
    const float LOG2E = 1.4426950408889634f;
    float m_ref = -INFINITY, l_run = 0.f;        // m_ref in log2 units, shared by the G threads of a row

    for (int j = 0; j < nblocks; ++j) {
      const int kvalid = Nb - j * BKEY - g * KT;        // valid keys among this thread's KT (may be <= 0 in the last block)
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      if (threadIdx.x == 0) astamp(a, 4 + 4 * j + 2);
      uint32_t v[KT];
      if (KT == 64) {
        tmem_ld32(tmem_Sg, reinterpret_cast<uint32_t(&)[32]>(v[0]));
        tmem_ld32(tmem_Sg + 32u, reinterpret_cast<uint32_t(&)[32]>(v[KT - 32]));
      } else {
        tmem_ld32(tmem_Sg, reinterpret_cast<uint32_t(&)[32]>(v[0]));
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      if (kvalid < KT) {                          // last, ragged block: keys beyond N do not exist
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if (i >= kvalid) v[i] = 0xff800000u;        // -inf
      }
      float mxa = -INFINITY, mxb = -INFINITY, mxc = -INFINITY, mxd = -INFINITY;        // four chains: the FMNMX3 latency is exposed otherwise
#pragma unroll
      for (int i = 0; i < KT; i += 8) {
        mxa = max3(mxa, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        mxb = max3(mxb, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        mxc = max3(mxc, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
        mxd = max3(mxd, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
      }
      float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
      // the row maximum over all key groups (group 0 always holds a valid key, so it is finite)
      float* xb = xch + (j & 1) * (4 * BQ);
      xb[g * BQ + r] = mx;
      rows_bar<G>(qd);
#pragma unroll
      for (int o = 1; o < G; ++o) mx = fmaxf(mx, xb[((g + o) & (G - 1)) * BQ + r]);
      mx *= LOG2E;
      // lazy running maximum
      const bool move = mx > m_ref + RESCALE_TAU;
      const float m_new = move ? mx : m_ref;
      const float alpha = move ? ex2(m_ref - m_new) : 1.0f;          // exp2(-inf) = 0 on the first block
      m_ref = m_new;
      l_run *= alpha;
      // P = exp2(s*log2e - m) -> 16 bit, kept in registers; fp32 row sum (packed accumulator)
      const unsigned long long sc2 = pk2(LOG2E, LOG2E), nm2 = pk2(-m_ref, -m_ref);
      unsigned long long sum2 = pk2(0.f, 0.f);
      uint32_t pkd[KT / 2];
#pragma unroll
      for (int u = 0; u < KT / 2; ++u) {
        const int i = u * 2;
        float e0, e1, p0, p1;
        up2(fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), e0, e1);
        if ((u & 3) < POLY) {
          ex2_fma_pair(e0, e1, p0, p1);
        } else {
          p0 = ex2(e0);
          p1 = ex2(e1);
        }
        sum2 = add2(sum2, pk2(p0, p1));
        pkd[u] = pk16<F16>(p0, p1);
      }
      float s0, s1;
      up2(sum2, s0, s1);
      l_run += s0 + s1;
      if (j > 0) {
        mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);                 // O += P_{j-1} V_{j-1} has landed (issued a whole block ago)
        tc_fence_after();
      }
      tmem_st_n(tmem_Pg, pkd);
      if (j > 0 && __any_sync(0xffffffffu, move)) {                  // rescale this thread's share of O (rare after the first blocks)
        uint32_t o[16];
#pragma unroll
        for (int cb = 0; cb < OC; cb += 16) {
          tmem_ld16(tmem_Og + (uint32_t)cb, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tmem_Og + (uint32_t)cb, o);
        }
      }
      tmem_st_wait();
      tc_fence_before();                 // order our TMEM writes before the MMA that reads P / accumulates into O
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (threadIdx.x == 0) astamp(a, 4 + 4 * j + 3);
    }
    // ---- row sum over all key groups, normalise, store ----
    float* xb = xch + (nblocks & 1) * (4 * BQ);          // the buffer the last block did not use
    xb[g * BQ + r] = l_run;
    rows_bar<G>(qd);
    float l_all = 0.f;
#pragma unroll
    for (int o = 0; o < G; ++o) l_all += xb[o * BQ + r];                 // the same order in every thread of the row
    const float scale = 1.0f / l_all;
    mbar_wait(pv_done, (uint32_t)(nblocks - 1) & 1u);
    tc_fence_after();
    if (threadIdx.x == 0) astamp(a, 2);
    uint32_t o[OC];
    tmem_ld_n(tmem_Og, o);
    tmem_ld_wait();
    const int q = q0 + r;
    if (q < Nb) {
      __nv_bfloat16* dst = a.out + ((a.seq_off ? (long)rb0 : (long)b * a.N) + q) * a.ldo + h * HD + g * OC;
#pragma unroll
      for (int k = 0; k < OC / 8; ++k) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = pk16<F16>(__uint_as_float(o[k * 8 + 2 * i]) * scale, __uint_as_float(o[k * 8 + 2 * i + 1]) * scale);
        *reinterpret_cast<uint4*>(dst + k * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }

  if (threadIdx.x == 0) astamp(a, 3);
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

}  // namespace

void attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vT, int ldv, __nv_bfloat16* out, int S, int N, int H, cudaStream_t stream, int f16,
                  const int* d_seq_off, const int* d_seq_len, long total_rows) {
  B2_CHECK(S > 0 && N > 0 && H > 0, "attention_tc: empty problem");
  B2_CHECK(ldv % 8 == 0 && ldv >= N, "attention_tc: V^T row stride must be a multiple of 8 and >= N");
  CUtensorMap map_qk, map_v;
  const bool ragged = d_seq_off != nullptr;
  B2_CHECK(ragged == (d_seq_len != nullptr) && (!ragged || total_rows > 0), "attention_tc: ragged batches need offsets, lengths and the row total");
  if (ragged) {
    // one concatenated [total_rows][2*H*64] tensor (rows past the end read as zero); N = the longest sequence
    tc_encode_map(&map_qk, qk, (uint64_t)2 * H * HD, (uint64_t)total_rows, 1, (uint64_t)2 * H * HD, (uint64_t)total_rows * 2 * H * HD, BQ);
    // V^T rows are ldv wide for every sequence: the keys beyond a sequence's length hold stale (finite) values and meet P = 0
    tc_encode_map(&map_v, vT, (uint64_t)ldv, (uint64_t)HD, (uint64_t)S * H, (uint64_t)ldv, (uint64_t)HD * ldv, HD);
  } else {
    // q|k: [S][N][2*H*64] -> dims {2*H*64, N, S}, box {64, 128, 1}
    tc_encode_map(&map_qk, qk, (uint64_t)2 * H * HD, (uint64_t)N, (uint64_t)S, (uint64_t)2 * H * HD, (uint64_t)N * 2 * H * HD, BQ);
    // V^T: [S*H][64][ldv] -> dims {N keys, 64 d, S*H}, box {64 keys, 64 d, 1}
    tc_encode_map(&map_v, vT, (uint64_t)N, (uint64_t)HD, (uint64_t)S * H, (uint64_t)ldv, (uint64_t)HD * ldv, HD);
  }
  // exponentials on the FMA pipes: B200TTS_ATTN_POLY = 0 .. 3 of every 4 score pairs (read once)
  static const int poly = [] {
    const char* e = getenv("B200TTS_ATTN_POLY");
    const int p = e ? atoi(e) : ATTN_POLY_DEFAULT;
    return p < 0 ? 0 : (p > 3 ? 3 : p);
  }();
  // softmax threads per query row: B200TTS_ATTN_G = 2 or 4 (read once)
  static const int gsel = [] {
    const char* e = getenv("B200TTS_ATTN_G");
    const int g = e ? atoi(e) : ATTN_G_DEFAULT;
    return g == 2 ? 0 : 1;
  }();
  using KernelFn = void (*)(CUtensorMap, CUtensorMap, AttnArgs);
  static const KernelFn kernels[2][2][4] = {
      {{attn_tc_kernel<false, 0, 2>, attn_tc_kernel<false, 1, 2>, attn_tc_kernel<false, 2, 2>, attn_tc_kernel<false, 3, 2>},
       {attn_tc_kernel<true, 0, 2>, attn_tc_kernel<true, 1, 2>, attn_tc_kernel<true, 2, 2>, attn_tc_kernel<true, 3, 2>}},
      {{attn_tc_kernel<false, 0, 4>, attn_tc_kernel<false, 1, 4>, attn_tc_kernel<false, 2, 4>, attn_tc_kernel<false, 3, 4>},
       {attn_tc_kernel<true, 0, 4>, attn_tc_kernel<true, 1, 4>, attn_tc_kernel<true, 2, 4>, attn_tc_kernel<true, 3, 4>}}};
  static PerDeviceOnce once;
  if (once.first())
    for (int g = 0; g < 2; ++g)
      for (int t = 0; t < 2; ++t)
        for (int p = 0; p < 4; ++p) B2_CUDA(cudaFuncSetAttribute(kernels[g][t][p], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const KernelFn kernel = kernels[gsel][f16 ? 1 : 0][poly];
  AttnArgs a{N, H, out, H * HD, d_seq_off, d_seq_len, nullptr};
  static DevBuf<unsigned long long> trace_buf;
  const char* trace_path = getenv("B200TTS_ATTN_TRACE");
  const size_t ncta = (size_t)ceil_div(N, BQ) * H * S;
  if (trace_path) {
    trace_buf.reserve(ncta * 64);
    B2_CUDA(cudaMemsetAsync(trace_buf.p, 0, ncta * 64 * 8, stream));
    a.trace = trace_buf.p;
  }
  B2_CHECK(S <= 65535, "attention_tc: too many sequences");
  dim3 grid(ceil_div(N, BQ), H, S);
  launch_pdl(kernel, grid, dim3(nthreads(gsel == 0 ? 2 : 4)), (size_t)SMEM_BYTES, stream, map_qk, map_v, a);
  B2_LAUNCH_CHECK();
  count_launch();
  if (trace_path) {                                        // debug only: dump the stamps of this launch
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      std::vector<unsigned long long> h(ncta * 64);
      B2_CUDA(cudaMemcpyAsync(h.data(), trace_buf.p, h.size() * 8, cudaMemcpyDeviceToHost, stream));
      B2_CUDA(cudaStreamSynchronize(stream));
      if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    }
  }
}

}  // namespace b200tts

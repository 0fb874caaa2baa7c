// Non-causal, unmasked, unscaled flash attention for the F5 DiT (modules.py:449-468: softmax(q @ k) @ v in
// fp32, the 1/sqrt(d) scale already folded into Wq/Wk at export) on tcgen05 tensor cores.
//
// One CTA = one (batch, head, 128-query tile). Per 128-key block j:
//   S_j  = Q K_j^T   tcgen05.mma  M=128 N=128 K=64   (Q, K tiles: TMA boxes of the [2N][2048] q|k tensor, SWIZZLE_128B)
//   P_j  = exp2(S_j log2e - m)   256 softmax threads: thread (row r, key half g) owns 64 keys of its row in ONE pass
//          (the 64 scores are pulled from TMEM into registers, which also frees the S buffer so that S_{j+1} is computed
//          while the exponentials run), packed FFMA2 / FADD2 arithmetic, P (bf16) to shared memory in the K-major
//          SWIZZLE_128B layout the next MMA wants
//   O_g += P_jg V_jg tcgen05.mma  M=128 N=64 K=64 per key half, accumulated IN TMEM across blocks.
// The two key halves of a row are independent split-KV streams: each has its own running maximum, row sum and O
// accumulator, so the two threads of a row never synchronise inside the loop; they are merged once at the end
// (O = (O_a 2^(m_a-m) + O_b 2^(m_b-m)) / (l_a 2^(m_a-m) + l_b 2^(m_b-m))) through shared memory.
// The running maximum is lazy (FlashAttention-4 style): it only moves when a block's maximum exceeds it by more than 8
// (log2 units), and only then is O_g rescaled in TMEM (tcgen05.ld / multiply / tcgen05.st, warp-voted); P stays below
// 2^8, which bf16 and the fp32 accumulators hold without loss.
// The first version (r01b: 30.5 us per call, 16.8 M instructions, issue-bound) walked S twice with 128 threads, kept O
// in registers and folded PV into it every block (64 FFMA + 2 TMEM loads per row per block).
// 320 threads: warps 0-7 = softmax/epilogue, warp 8 = TMA producer, warp 9 = MMA issuer + TMEM owner.
// Shared memory 112 KB and 256 TMEM columns per CTA -> two CTAs per SM overlap each other's softmax and MMA phases.
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
constexpr int NTHREADS = 320;
constexpr int WARP_TMA = 8, WARP_MMA = 9;       // highest warp ids: the SMSP arbiter favours them over the softmax warps
constexpr int Q_BYTES = BQ * HD * 2;            // 16 KB
constexpr int K_BYTES = BKEY * HD * 2;          // 16 KB
constexpr int V_BYTES = HD * BKEY * 2;          // 16 KB = two [64 d][64 keys] chunks
constexpr int P_BYTES = BQ * BKEY * 2;          // 32 KB = two [128 rows][64 keys] chunks
constexpr int KV_STAGES = 2;
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + P_BYTES + 128;      // x2 CTAs + 2 KB reserved <= 228 KB
constexpr float RESCALE_TAU = 8.0f;             // log2 units

struct AttnArgs {
  int N, H;
  __nv_bfloat16* out;
  int ldo;
  int f16;            // q, k, v, P and out are IEEE fp16 instead of bf16
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

__device__ __forceinline__ uint32_t pk16(float x, float y, int half) {
  if (half) { __half2 h = __floats2half2_rn(x, y); return *reinterpret_cast<uint32_t*>(&h); }
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
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(NTHREADS, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap map_qk,
                                                              const __grid_constant__ CUtensorMap map_v, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KV_STAGES * K_BYTES;
  uint8_t* sP = sV + KV_STAGES * V_BYTES;
  float* stat = reinterpret_cast<float*>(sQ);                      // [2 halves][m, l][128 rows]: reuses the Q tile once the
                                                                   // last S MMA has completed (pv_done of the last block)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                  // [2]  K and V stages are released separately: a K tile is dead as soon as
  uint64_t* k_empty = bars + 3;                 // [2]  S_j = Q K_j^T has run, its V tile only after O += P_j V_j, a whole
  uint64_t* v_full = bars + 5;                  // [2]  softmax later. With one barrier pair K_{j+1} arrived too late and the
  uint64_t* v_empty = bars + 7;                 // [2]  softmax threads spent 21 % of their samples waiting for S (ncu r01p).
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;                 // softmax threads hold S in registers: the S buffer may be overwritten
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

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
    mbar_init(s_free, 256);
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();                                          // after the TMEM allocation (common.cuh)
  pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) astamp(a, 0);
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;     // O_a: +128..191, O_b: +192..255

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
      const uint32_t fmt = a.f16 ? 0u : ((1u << 7) | (1u << 10));        // kind::f16 operand format: 1 = bf16, 0 = fp16
      const uint32_t idesc_s = (1u << 4) | fmt | ((uint32_t)(BKEY >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | fmt | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint64_t dQ = make_desc_sw128(smem_u32(sQ));
      const uint64_t dP0 = make_desc_sw128(smem_u32(sP)), dP1 = make_desc_sw128(smem_u32(sP + P_BYTES / 2));
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
        for (int k = 0; k < 4; ++k) {                               // the two key halves are independent accumulators
          umma_bf16(tmem_O, dP0 + (uint64_t)(2 * k), dV0 + (uint64_t)(2 * k), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_O + 64, dP1 + (uint64_t)(2 * k), dV1 + (uint64_t)(2 * k), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(pv_done);
        umma_commit(&v_empty[s]);
        astamp(a, 4 + 4 * j + 1);
      }
    }
  } else {
    // ===== softmax + epilogue: row r of the query tile, key half g of every block =====
    const int qd = warp & 3, g = warp >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_Og = tmem_O + (uint32_t)(g * 64);
    const float LOG2E = 1.4426950408889634f;
    float m_ref = -INFINITY, l_run = 0.f;        // m_ref in log2 units
    uint8_t* prow = sP + g * (P_BYTES / 2) + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;

    for (int j = 0; j < nblocks; ++j) {
      const int kvalid = Nb - j * BKEY - g * 64;        // valid keys among this thread's 64 (may be <= 0 in the last block)
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      if (threadIdx.x == 0) astamp(a, 4 + 4 * j + 2);
      uint32_t v[2][32];
      tmem_ld32(tmem_S + lane_addr + (uint32_t)(g * 64), v[0]);
      tmem_ld32(tmem_S + lane_addr + (uint32_t)(g * 64 + 32), v[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);
      if (kvalid < 64) {                          // last, ragged block: keys beyond N do not exist
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= kvalid) v[c][i] = 0xff800000u;        // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]));
      mx *= LOG2E;
      // lazy running maximum
      const bool move = mx > m_ref + RESCALE_TAU;
      const float m_new = move ? mx : m_ref;
      const float alpha = move ? ex2(m_ref - m_new) : 1.0f;          // exp2(-inf) = 0 on the first block
      m_ref = m_new;
      l_run *= alpha;
      if (j > 0) {
        mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);                 // O += P_{j-1} V_{j-1} has landed; P buffer is free again
        tc_fence_after();
        if (__any_sync(0xffffffffu, move)) {                         // rescale O_g in TMEM (rare after the first blocks)
#pragma unroll 1
          for (int cb = 0; cb < HD; cb += 16) {                      // 16 columns at a time: the 64 scores stay live
            uint32_t o[16];
            tmem_ld16(tmem_Og + lane_addr + (uint32_t)cb, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_Og + lane_addr + (uint32_t)cb, o);
          }
          tmem_st_wait();
        }
      }
      // P = exp2(s*log2e - m) -> bf16 -> swizzled smem; fp32 row sum (packed accumulator). No key seen yet (m = -inf) -> P = 0
      const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
      const unsigned long long sc2 = pk2(LOG2E, LOG2E), nm2 = pk2(-m_use, -m_use);
      unsigned long long sum2 = pk2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        // 32 keys = four 16-byte chunks of this row's 128-byte (64-key) line: chunk index c*4 + q, swizzled by the row
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          uint32_t pkd[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = qq * 8 + u * 2;
            float e0, e1;
            up2(fma2(pk2(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])), sc2, nm2), e0, e1);
            const float p0 = ex2(e0), p1 = ex2(e1);
            sum2 = add2(sum2, pk2(p0, p1));
            pkd[u] = pk16(p0, p1, a.f16);
          }
          const int chunk = (c * 4 + qq) ^ sw;
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pkd[0], pkd[1], pkd[2], pkd[3]);
        }
      }
      float s0, s1;
      up2(sum2, s0, s1);
      l_run += s0 + s1;
      tc_fence_before();                 // order our TMEM accesses before the MMA warp's next writes
      fence_proxy_async();               // make the generic-proxy smem writes of P visible to the tensor core
      mbar_arrive(p_full);
      if (threadIdx.x == 0) astamp(a, 4 + 4 * j + 3);
    }
    // ---- merge the two key halves of each row, normalise, store ----
    mbar_wait(pv_done, (uint32_t)(nblocks - 1) & 1u);
    tc_fence_after();
    if (threadIdx.x == 0) astamp(a, 2);
    stat[(g * 2 + 0) * BQ + r] = m_ref;
    stat[(g * 2 + 1) * BQ + r] = l_run;
    softmax_bar();
    const float m_o = stat[((g ^ 1) * 2 + 0) * BQ + r], l_o = stat[((g ^ 1) * 2 + 1) * BQ + r];
    const float m_all = fmaxf(m_ref, m_o);       // finite: key half 0 always holds at least one valid key
    const float w_me = m_ref == -INFINITY ? 0.f : ex2(m_ref - m_all);
    const float w_ot = m_o == -INFINITY ? 0.f : ex2(m_o - m_all);
    const float scale = w_me / (l_run * w_me + l_o * w_ot);
    // thread (r, g) stores output columns [32g, 32g+32); the other 32 columns of its O_g go to the partner through the
    // (now idle) P buffer: [2 halves][128 rows][32 floats], 16-byte chunks swizzled by the row
    float* xch = reinterpret_cast<float*>(sP);
    uint32_t mine[32];
    {
      uint32_t o[32];
      tmem_ld32(tmem_Og + lane_addr + (uint32_t)((g ^ 1) * 32), o);        // the partner's columns
      tmem_ld32(tmem_Og + lane_addr + (uint32_t)(g * 32), mine);
      tmem_ld_wait();
      float* dstx = xch + ((g ^ 1) * BQ + r) * 32;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(dstx + ((k ^ sw) << 2)) =
            make_float4(__uint_as_float(o[k * 4]) * scale, __uint_as_float(o[k * 4 + 1]) * scale, __uint_as_float(o[k * 4 + 2]) * scale,
                        __uint_as_float(o[k * 4 + 3]) * scale);
    }
    softmax_bar();
    const int q = q0 + r;
    if (q < Nb) {
      const float* srcx = xch + (g * BQ + r) * 32;
      __nv_bfloat16* dst = a.out + ((a.seq_off ? (long)rb0 : (long)b * a.N) + q) * a.ldo + h * HD + g * 32;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const float4 x0 = *reinterpret_cast<const float4*>(srcx + ((k ^ sw) << 2));
        const float4 x1 = *reinterpret_cast<const float4*>(srcx + (((k + 1) ^ sw) << 2));
        uint32_t w[4];
        w[0] = pk16(__uint_as_float(mine[k * 4 + 0]) * scale + x0.x, __uint_as_float(mine[k * 4 + 1]) * scale + x0.y, a.f16);
        w[1] = pk16(__uint_as_float(mine[k * 4 + 2]) * scale + x0.z, __uint_as_float(mine[k * 4 + 3]) * scale + x0.w, a.f16);
        w[2] = pk16(__uint_as_float(mine[k * 4 + 4]) * scale + x1.x, __uint_as_float(mine[k * 4 + 5]) * scale + x1.y, a.f16);
        w[3] = pk16(__uint_as_float(mine[k * 4 + 6]) * scale + x1.z, __uint_as_float(mine[k * 4 + 7]) * scale + x1.w, a.f16);
        *reinterpret_cast<uint4*>(dst + k * 4) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }

  if (threadIdx.x == 0) astamp(a, 3);
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
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
  static std::once_flag once;
  std::call_once(once, [] {
    B2_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  });
  AttnArgs a{N, H, out, H * HD, f16, d_seq_off, d_seq_len, nullptr};
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
  launch_pdl(attn_tc_kernel, grid, dim3(NTHREADS), (size_t)SMEM_BYTES, stream, map_qk, map_v, a);
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

// Non-causal, unmasked, unscaled flash attention for the F5 DiT (modules.py:449-468: softmax(q @ k) @ v in
// fp32, the 1/sqrt(d) scale already folded into Wq/Wk at export) on tcgen05 tensor cores.
//
// One CTA = one (batch, head, 128-query tile). Per 128-key block:
//   S  = Q K^T   tcgen05.mma  M=128 N=128 K=64   (Q, K tiles: TMA boxes of the [2N][2048] q|k tensor, SWIZZLE_128B)
//   P  = exp2((S - m) log2e)   128 softmax threads, one row each, read S from TMEM, online max/sum in fp32,
//        write P (bf16) to shared memory in the K-major SWIZZLE_128B layout the next MMA wants
//   PV = P V     tcgen05.mma  M=128 N=64 K=128  (V^T tiles [64 d][keys] from the transposed-V buffer the QKV GEMM
//        epilogue wrote) into a second TMEM buffer; the softmax threads fold it into register accumulators
//        O = O * alpha + PV, so no TMEM read-modify-write correction pass is needed.
// 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = softmax/epilogue.
// Shared memory 112 KB and 256 TMEM columns per CTA -> two CTAs per SM overlap each other's softmax and MMA phases.
#include "attention_tc.cuh"

#include <mutex>

#include "rowgemm_tc.cuh"
#include "tc_ptx.cuh"

namespace b200tts {

namespace {

using namespace tc;

constexpr int BQ = 128, BKEY = 128, HD = 64;
constexpr int NTHREADS = 192;
constexpr int Q_BYTES = BQ * HD * 2;            // 16 KB
constexpr int K_BYTES = BKEY * HD * 2;          // 16 KB
constexpr int V_BYTES = HD * BKEY * 2;          // 16 KB = two [64 d][64 keys] chunks
constexpr int P_BYTES = BQ * BKEY * 2;          // 32 KB = two [128 rows][64 keys] chunks
constexpr int KV_STAGES = 2;
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + P_BYTES + 128;

struct AttnArgs {
  int N, H;
  __nv_bfloat16* out;
  int ldo;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(NTHREADS, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap map_qk,
                                                              const __grid_constant__ CUtensorMap map_v, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KV_STAGES * K_BYTES;
  uint8_t* sP = sV + KV_STAGES * V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                 // [2]
  uint64_t* kv_empty = bars + 3;                // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* pv_full = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nblocks = (a.N + BKEY - 1) / BKEY;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();       // SWIZZLE_128B tiles need a 1024-byte aligned base
    prefetch_tmap(&map_qk);
    prefetch_tmap(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_PV = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_3d(sQ, &map_qk, q_full, h * HD, q0, b);
      const int bh = b * a.H + h;
      for (int j = 0; j < nblocks; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (uint32_t)(j / KV_STAGES) & 1u;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        mbar_expect_tx(&kv_full[s], K_BYTES + V_BYTES);
        tma_load_3d(sK + s * K_BYTES, &map_qk, &kv_full[s], a.H * HD + h * HD, j * BKEY, b);
        tma_load_3d(sV + s * V_BYTES, &map_v, &kv_full[s], j * BKEY, 0, bh);
        tma_load_3d(sV + s * V_BYTES + V_BYTES / 2, &map_v, &kv_full[s], j * BKEY + 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKEY >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint64_t dQ = make_desc_sw128(smem_u32(sQ));
      const uint64_t dP0 = make_desc_sw128(smem_u32(sP)), dP1 = make_desc_sw128(smem_u32(sP + P_BYTES / 2));
      mbar_wait(q_full, 0);
      for (int j = 0; j < nblocks; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (uint32_t)(j / KV_STAGES) & 1u;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        const uint64_t dK = make_desc_sw128(smem_u32(sK + s * K_BYTES));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_S, dQ + (uint64_t)(2 * k), dK + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_full, (uint32_t)j & 1u);
        tc_fence_after();
        const uint64_t dV0 = make_desc_sw128(smem_u32(sV + s * V_BYTES));
        const uint64_t dV1 = make_desc_sw128(smem_u32(sV + s * V_BYTES + V_BYTES / 2));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_PV, dP0 + (uint64_t)(2 * k), dV0 + (uint64_t)(2 * k), idesc_pv, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_PV, dP1 + (uint64_t)(2 * k), dV1 + (uint64_t)(2 * k), idesc_pv, 1u);
        umma_commit(pv_full);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ===== softmax + epilogue: row r of the query tile =====
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;

    for (int j = 0; j < nblocks; ++j) {
      const int kvalid = a.N - j * BKEY;          // >= 1
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int cb = 0; cb < BKEY; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = __uint_as_float(v[i]);
          mx = (cb + i < kvalid) ? fmaxf(mx, sv) : mx;
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2((m_run - m_new) * LOG2E);     // exp2(-inf) = 0 on the first block
      const float mb = m_new * LOG2E;
      // pass 2: P = exp2(s*log2e - m*log2e) -> bf16 -> swizzled smem; fp32 row sum
      float sum = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < BKEY; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)cb, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (cb + i < kvalid) ? ex2(fmaf(__uint_as_float(v[i]), LOG2E, -mb)) : 0.f;
          float p1 = (cb + i + 1 < kvalid) ? ex2(fmaf(__uint_as_float(v[i + 1]), LOG2E, -mb)) : 0.f;
          sum += p0 + p1;
          __nv_bfloat162 pp = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&pp);
        }
        // 32 keys = four 16-byte chunks; chunk index within the 64-key (128 B) row: (cb % 64) / 8 + q
        uint8_t* base = prow + (cb >> 6) * (P_BYTES / 2);
        const int c0 = (cb & 63) >> 3;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int chunk = (c0 + qq) ^ sw;
          *reinterpret_cast<uint4*>(base + chunk * 16) = make_uint4(pk[qq * 4 + 0], pk[qq * 4 + 1], pk[qq * 4 + 2], pk[qq * 4 + 3]);
        }
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      tc_fence_before();                 // order our TMEM reads of S before the MMA warp's next write to it
      fence_proxy_async();               // make the generic-proxy smem writes of P visible to the tensor core
      mbar_arrive(p_full);
      // fold PV into the register accumulator
      mbar_wait(pv_full, (uint32_t)j & 1u);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < HD; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_PV + lane_addr + (uint32_t)cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[cb + i] = fmaf(o[cb + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
    }
    const int q = q0 + r;
    if (q < a.N) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = a.out + ((long)b * a.N + q) * a.ldo + h * HD;
#pragma unroll
      for (int i = 0; i < HD; i += 8) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __nv_bfloat162 pp = __floats2bfloat162_rn(o[i + 2 * k] * inv, o[i + 2 * k + 1] * inv);
          w[k] = *reinterpret_cast<uint32_t*>(&pp);
        }
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

void attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vT, int ldv, __nv_bfloat16* out, int N, int H, cudaStream_t stream) {
  B2_CHECK(N > 0 && H > 0, "attention_tc: empty problem");
  B2_CHECK(ldv % 8 == 0 && ldv >= N, "attention_tc: V^T row stride must be a multiple of 8 and >= N");
  CUtensorMap map_qk, map_v;
  // q|k: [2][N][2*H*64] bf16 -> dims {2*H*64, N, 2}, box {64, 128, 1}
  tc_encode_map(&map_qk, qk, (uint64_t)2 * H * HD, (uint64_t)N, 2, (uint64_t)2 * H * HD, (uint64_t)N * 2 * H * HD, BQ);
  // V^T: [2*H][64][ldv] bf16 -> dims {N keys, 64 d, 2H}, box {64 keys, 64 d, 1}
  tc_encode_map(&map_v, vT, (uint64_t)N, (uint64_t)HD, (uint64_t)2 * H, (uint64_t)ldv, (uint64_t)HD * ldv, HD);
  static std::once_flag once;
  std::call_once(once, [] {
    B2_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  });
  AttnArgs a{N, H, out, H * HD};
  dim3 grid(ceil_div(N, BQ), H, 2);
  attn_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(map_qk, map_v, a);
  B2_LAUNCH_CHECK();
  count_launch();
}

}  // namespace b200tts

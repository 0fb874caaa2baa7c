// BigVGAN-v2 generator, channels-last, B200-native.
//
// Reference semantics (quirks included): BigVGAN/Export_BigVGAN.py:37-49 (x32767, clamp, truncating cast,
// tanh forced on), BigVGAN/modeling_modified/bigvgan.py:384-410 (forward), :132-140 (AMPBlock1),
// :359-382 (pad tables; index -1 = 15-sample pads -> +30 output samples).
//
// Data layout: every activation is (B, L, C) with C contiguous, so that
//   * each Conv1d / ConvTranspose1d is a shifted-row GEMM (rowgemm.cuh) whose A tiles are plain
//     2-D boxes of the activation tensor (TMA-able, zero fill at the sequence ends = zero padding),
//   * the anti-aliased activation streams coalesced rows (aa_act.cu),
//   * bias / residual add / MRF mean (x 1/3) live in the GEMM epilogue.
#include "bigvgan.cuh"

#include "aa_act.cuh"
#include "f5_kernels.cuh"
#include "layout.cuh"
#include "rowgemm.cuh"
#include "rowgemm_tc.cuh"

namespace b200tts {

namespace {

// W (Cout, Cin, k) -> out[j][c][n]
__global__ void prep_conv_w_kernel(const float* __restrict__ W, float* __restrict__ out, int Cout, int Cin, int k) {
  const long total = (long)Cout * Cin * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Cout);
    const int c = (int)((i / Cout) % Cin);
    const int j = (int)(i / ((long)Cout * Cin));
    out[i] = W[((long)n * Cin + c) * k + j];
  }
}

// ConvTranspose1d W (Cin, Cout, k), k = taps*u:
//   taps = 2 (k = 2u, padding u/2): out[tap][c][r*Cout + n] = W[c][n][tap == 0 ? r + u : r]
//   taps = 1 (k = u,  padding 0)  : out[0][c][r*Cout + n]   = W[c][n][r]          (no overlap between input samples)
__global__ void prep_convtr_w_kernel(const float* __restrict__ W, float* __restrict__ out, int Cin, int Cout, int u, int taps) {
  const long N = (long)u * Cout;
  const long total = (long)taps * Cin * N;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i % N);
    const int c = (int)((i / N) % Cin);
    const int tap = (int)(i / (N * Cin));
    const int r = col / Cout, n = col % Cout;
    const int j = (taps == 2 && tap == 0) ? r + u : r;
    out[i] = W[((long)c * Cout + n) * (taps * u) + j];
  }
}

// bias_out[r*C + n] = bias[n] + cond[n]  (r < reps): the per-call conditioning vector folded into a GEMM bias
__global__ void cond_bias_kernel(const float* __restrict__ bias, const float* __restrict__ cond, float* __restrict__ out, int C, int reps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * reps) out[i] = bias[i % C] + (cond ? cond[i % C] : 0.f);
}

__global__ void replicate_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int Cout, int u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Cout * u) out[i] = b[i % Cout];
}

__global__ void snake_params_kernel(const float* __restrict__ alpha_log, const float* __restrict__ beta_log,
                                    float* __restrict__ alpha, float* __restrict__ inv_beta, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    alpha[i] = expf(alpha_log[i]);
    inv_beta[i] = 1.0f / (expf(beta_log[i]) + 1e-9f);
  }
}

// conv_post (k=7, pad 3, C -> 1, no bias) + tanh + x32767 + clamp + truncating int16 cast.
// x: (B, L, C) channels-last, so the 7xC window of one output sample is contiguous in memory.
template <int C>
__global__ void __launch_bounds__(256) post_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, float bias,
                                                        int L, int16_t* __restrict__ pcm, float* __restrict__ wave) {
  __shared__ float ws[7 * C];
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float* xb = x + (long)b * L * C;
  float acc = bias;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int tt = t + j - 3;
    if (tt < 0 || tt >= L) continue;
    const float4* row = reinterpret_cast<const float4*>(xb + (long)tt * C);
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = __ldg(row + c4);
      acc = fmaf(v.x, ws[j * C + c4 * 4 + 0], acc);
      acc = fmaf(v.y, ws[j * C + c4 * 4 + 1], acc);
      acc = fmaf(v.z, ws[j * C + c4 * 4 + 2], acc);
      acc = fmaf(v.w, ws[j * C + c4 * 4 + 3], acc);
    }
  }
  float v = tanhf(acc) * 32767.0f;
  v = fminf(fmaxf(v, -32768.0f), 32767.0f);
  pcm[(long)b * L + t] = (int16_t)v;   // float -> int conversion truncates toward zero, like torch .to(int16)
  if (wave) wave[(long)b * L + t] = v;
}

struct ConvW {
  DevBuf<float> w;      // fp32 [taps][Cin][N]
  DevBuf<float> bias;   // [N]
  TcWeight tc[2];       // 16-bit [taps][N][Cin] (fast path), built lazily per operand type: [0] bf16, [1] fp16
  int Cin = 0, N = 0, taps = 0;
};

struct Snake {
  DevBuf<float> alpha, inv_beta;
};

struct Stage {
  int u = 0, Cin = 0, C = 0;
  ConvW up;                         // as a 2-tap (k = 2u) or 1-tap (k = u) rowgemm with N = u*C
  DevBuf<float> up_bias_c;          // u*C: bias + per-call conditioning vector (IndexTTS)
  ConvW c1[3][3], c2[3][3];         // [resblock][m]
  Snake act[3][6];
  int k[3] = {0, 0, 0};
  int dil[3][3];
};

}  // namespace

struct BigVGANModel {
  int n_mels = 0, C0 = 0, nstages = 0, hop = 1;
  std::string prefix;               // "bigvgan." (mel -> PCM session) or "ivgan." (IndexTTS_F: GPT latent -> PCM)
  // IndexTTS_F extras (IndexTTS/Export_IndexTTS.py:300-314): final LayerNorm of the GPT latent, conditioning adds, conv_post bias
  DevBuf<float> fn_w, fn_b;         // gpt.final_norm (empty for the mel vocoder)
  DevBuf<float> pre_bias_c;         // conv_pre bias + cond_layer vector
  DevBuf<float> latent;             // normalised latent (S-2, gpt_dim)
  float post_bias = 0.f;
  ConvW pre;
  std::vector<Stage> stages;
  Snake post_act;
  DevBuf<float> post_w;             // [7][Clast]
  int Clast = 0;
  // workspaces (grow-only)
  DevBuf<float> mel_cl, xu, xs, abuf, cbuf;
  DevBuf<float> xa[3], xb[3];                      // per resblock branch (the three branches of a stage run concurrently)
  DevBuf<__nv_bfloat16> abuf16[3], cbuf16[3], mel16, xs16;
};

namespace {

void prep_conv(Engine& e, const std::string& wname, const std::string& bname, ConvW& cw) {
  const Tensor& W = e.weight(wname);
  B2_CHECK(W.shape.size() == 3, wname + " must be rank 3");
  const int Cout = (int)W.shape[0], Cin = (int)W.shape[1], k = (int)W.shape[2];
  cw.Cin = Cin; cw.N = Cout; cw.taps = k;
  cw.w.alloc((size_t)Cout * Cin * k);
  prep_conv_w_kernel<<<ceil_div((long)Cout * Cin * k, 256), 256, 0, e.stream>>>(W.data.p, cw.w.p, Cout, Cin, k);
  B2_LAUNCH_CHECK();
  if (!bname.empty()) {
    const Tensor& bt = e.weight(bname);
    B2_CHECK(bt.numel() == Cout, bname + " size mismatch");
    cw.bias.alloc(Cout);
    B2_CUDA(cudaMemcpyAsync(cw.bias.p, bt.data.p, Cout * sizeof(float), cudaMemcpyDeviceToDevice, e.stream));
  }
}

void prep_snake(Engine& e, const std::string& prefix, int C, Snake& s) {
  const Tensor& a = e.weight(prefix + ".alpha");
  const Tensor& b = e.weight(prefix + ".beta");
  B2_CHECK(a.numel() == C && b.numel() == C, prefix + " snake parameter size mismatch");
  s.alpha.alloc(C); s.inv_beta.alloc(C);
  snake_params_kernel<<<ceil_div(C, 128), 128, 0, e.stream>>>(a.data.p, b.data.p, s.alpha.p, s.inv_beta.p, C);
  B2_LAUNCH_CHECK();
}

}  // namespace

BigVGANModel* bigvgan_build(Engine& e, const std::string& P) {
  std::unique_ptr<BigVGANModel> m(new BigVGANModel());
  m->prefix = P;
  {
    const Tensor& f = e.weight(P + "aa_filter");
    B2_CHECK(f.numel() == 12, P + "aa_filter must hold 12 taps");
    float taps[12];
    B2_CUDA(cudaMemcpy(taps, f.data.p, sizeof(taps), cudaMemcpyDeviceToHost));
    aa_set_filter(taps);
  }
  prep_conv(e, P + "conv_pre.weight", P + "conv_pre.bias", m->pre);
  m->n_mels = m->pre.Cin; m->C0 = m->pre.N;
  B2_CHECK(m->n_mels % 4 == 0, "n_mels must be a multiple of 4");
  int ns = 0;
  while (e.has_weight(P + "ups." + std::to_string(ns) + ".0.weight")) ++ns;
  B2_CHECK(ns > 0, "no bigvgan.ups.* tensors loaded");
  m->nstages = ns;
  m->stages.resize(ns);
  int C = m->C0;
  for (int i = 0; i < ns; ++i) {
    Stage& st = m->stages[i];
    const Tensor& W = e.weight(P + "ups." + std::to_string(i) + ".0.weight");
    B2_CHECK(W.shape.size() == 3 && W.shape[0] == C, "ups weight shape");
    st.Cin = C; st.C = (int)W.shape[1];
    const int k = (int)W.shape[2];
    // stride: k = 2u with padding u/2 (upsample_kernel_sizes = 2 * upsample_rates, the 24 kHz v2 config), or k = u with
    // padding 0 (IndexTTS' later stages). The state dict does not carry the stride; "<prefix>upsample_rates" does when given.
    int u = k / 2;
    if (e.has_weight(P + "upsample_rates")) {
      const Tensor& R = e.weight(P + "upsample_rates");
      B2_CHECK(R.numel() >= ns, "upsample_rates shorter than the number of stages");
      float ru = 0.f;
      B2_CUDA(cudaMemcpy(&ru, R.data.p + i, sizeof(float), cudaMemcpyDeviceToHost));
      u = (int)(ru + 0.5f);
    }
    B2_CHECK(u > 0 && (k == 2 * u || k == u), "ConvTranspose1d kernel must be the stride or twice the stride");
    const int up_taps = k / u;
    st.u = u;
    B2_CHECK(up_taps == 1 || st.u % 2 == 0, "upsample rate must be even when kernel = 2*stride");
    m->hop *= st.u;
    const long N = (long)st.u * st.C;
    st.up.Cin = st.Cin; st.up.N = (int)N; st.up.taps = up_taps;
    st.up.w.alloc((size_t)up_taps * st.Cin * N);
    prep_convtr_w_kernel<<<ceil_div((long)up_taps * st.Cin * N, 256), 256, 0, e.stream>>>(W.data.p, st.up.w.p, st.Cin, st.C, st.u, up_taps);
    B2_LAUNCH_CHECK();
    const Tensor& bt = e.weight(P + "ups." + std::to_string(i) + ".0.bias");
    st.up.bias.alloc(N);
    replicate_bias_kernel<<<ceil_div(N, 128), 128, 0, e.stream>>>(bt.data.p, st.up.bias.p, st.C, st.u);
    B2_LAUNCH_CHECK();
    for (int j = 0; j < 3; ++j) {
      const std::string rp = P + "resblocks." + std::to_string(i * 3 + j) + ".";
      for (int mm = 0; mm < 3; ++mm) {
        prep_conv(e, rp + "convs1." + std::to_string(mm) + ".weight", rp + "convs1." + std::to_string(mm) + ".bias", st.c1[j][mm]);
        prep_conv(e, rp + "convs2." + std::to_string(mm) + ".weight", rp + "convs2." + std::to_string(mm) + ".bias", st.c2[j][mm]);
        B2_CHECK(st.c1[j][mm].Cin == st.C && st.c1[j][mm].N == st.C, "resblock conv shape");
        st.dil[j][mm] = 2 * mm + 1;     // resblock_dilation_sizes = (1, 3, 5) for every kernel size
      }
      st.k[j] = st.c1[j][0].taps;
      for (int a = 0; a < 6; ++a) prep_snake(e, rp + "activations." + std::to_string(a) + ".act", st.C, st.act[j][a]);
    }
    C = st.C;
  }
  m->Clast = C;
  prep_snake(e, P + "activation_post.act", C, m->post_act);
  {
    const Tensor& W = e.weight(P + "conv_post.weight");   // (1, C, 7); no bias in the 24 kHz v2 config, one in IndexTTS'
    B2_CHECK(W.shape.size() == 3 && W.shape[0] == 1 && W.shape[1] == C && W.shape[2] == 7, "conv_post shape");
    if (e.has_weight(P + "conv_post.bias")) {
      const Tensor& pb = e.weight(P + "conv_post.bias");
      B2_CHECK(pb.numel() == 1, "conv_post bias must be a scalar");
      B2_CUDA(cudaMemcpy(&m->post_bias, pb.data.p, sizeof(float), cudaMemcpyDeviceToHost));
    }
    m->post_w.alloc(7 * C);
    prep_conv_w_kernel<<<1, 256, 0, e.stream>>>(W.data.p, m->post_w.p, 1, C, 7);   // -> [j][c][0]
    B2_LAUNCH_CHECK();
  }
  if (e.has_weight(P + "final_norm.weight")) {           // IndexTTS_F: gpt.final_norm on the latent rows
    const Tensor& fw = e.weight(P + "final_norm.weight");
    const Tensor& fb = e.weight(P + "final_norm.bias");
    B2_CHECK(fw.numel() == m->n_mels && fb.numel() == m->n_mels, "final_norm size must equal the conv_pre input width");
    m->fn_w.alloc(m->n_mels); m->fn_b.alloc(m->n_mels);
    B2_CUDA(cudaMemcpyAsync(m->fn_w.p, fw.data.p, m->n_mels * sizeof(float), cudaMemcpyDeviceToDevice, e.stream));
    B2_CUDA(cudaMemcpyAsync(m->fn_b.p, fb.data.p, m->n_mels * sizeof(float), cudaMemcpyDeviceToDevice, e.stream));
  }
  m->pre_bias_c.alloc(m->C0);
  for (auto& st : m->stages) st.up_bias_c.alloc((size_t)st.u * st.C);
  B2_CUDA(cudaStreamSynchronize(e.stream));
  return m.release();
}

void bigvgan_free(BigVGANModel* m) { delete m; }

int bigvgan_num_mels(const BigVGANModel& m) { return m.n_mels; }
int bigvgan_num_stages(const BigVGANModel& m) { return m.nstages; }
int bigvgan_stage_channels(const BigVGANModel& m, int i) { return i < 0 ? m.C0 : m.stages[i].C; }
long bigvgan_out_samples(const BigVGANModel& m, int T) { return (long)m.hop * T + 30; }

namespace {

// fp32 [j][c][n] (the SIMT layout) -> fp32 [j][n][c] (K-major B operand), then cast to bf16 + TMA map
__global__ void jcn_to_jnc_kernel(const float* __restrict__ in, float* __restrict__ out, int taps, int Cin, int N) {
  const long total = (long)taps * Cin * N;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int n = (int)((i / Cin) % N);
    const int j = (int)(i / ((long)Cin * N));
    out[i] = in[((long)j * Cin + c) * N + n];
  }
}

void prep_tc(Engine& e, ConvW& cw, DevBuf<float>& tmp, int f16) {
  if (cw.tc[f16].ready) return;
  const long total = (long)cw.taps * cw.Cin * cw.N;
  tmp.reserve(total);
  jcn_to_jnc_kernel<<<ceil_div(total, 256), 256, 0, e.stream>>>(cw.w.p, tmp.p, cw.taps, cw.Cin, cw.N);
  B2_LAUNCH_CHECK();
  tc_weight_from_f32(cw.tc[f16], tmp.p, 1, cw.taps, cw.N, cw.Cin, e.stream, f16);
}

}  // namespace

void bigvgan_tc_prepare(Engine& e, BigVGANModel& m, int f16) {
  if (m.pre.tc[f16].ready) return;
  DevBuf<float> tmp;
  for (auto& st : m.stages) {
    prep_tc(e, st.up, tmp, f16);
    for (int j = 0; j < 3; ++j)
      for (int mm = 0; mm < 3; ++mm) { prep_tc(e, st.c1[j][mm], tmp, f16); prep_tc(e, st.c2[j][mm], tmp, f16); }
  }
  prep_tc(e, m.pre, tmp, f16);
  B2_CUDA(cudaStreamSynchronize(e.stream));
}

namespace {

struct Ctx {
  Engine& e;
  BigVGANModel& m;
  int B;
  bool fast;
  int f16;      // fast path operand type: 0 = bf16, 1 = fp16
  int t16;      // its type code as an output (rowgemm.cuh out_bf16 / aa_snake): 1 = bf16, 2 = fp16
  cudaStream_t s;   // the stream this (branch of the) forward pass enqueues on
};

// y = conv(x) with the shifted-row GEMM; x, out are (B, L, C) fp32 (or bf16 for the tc path's A operand)
void run_conv(Ctx& c, const char* tag, const ConvW& cw, const void* x, int L, int dil, void* out, int out_bf16,
              const float* res, int accumulate, float scale, int ldx = 0, const float* bias = nullptr,
              __nv_bfloat16* out2 = nullptr) {
  RowGemm p;
  if (ldx == 0) ldx = cw.Cin;
  p.x = x; p.x_bstride = (long)L * ldx; p.ldx = ldx; p.Lin = L;
  p.Cin = cw.Cin; p.N = cw.N; p.taps = cw.taps; p.dil = dil; p.center = (cw.taps - 1) / 2;
  p.M = L; p.B = c.B;
  p.out = out; p.o_bstride = (long)L * cw.N; p.ldo = cw.N; p.out_bf16 = out_bf16;
  p.bias = bias ? bias : cw.bias.p; p.res = res; p.accumulate = accumulate; p.scale = scale; p.out2 = c.fast ? out2 : nullptr;
  p.f16 = c.f16;
  ProfScope ps(c.e.prof, tag, c.s);
  if (c.fast) {
    rowgemm_tc(p, cw.tc[c.f16], c.s);
  } else {
    p.w = cw.w.p; p.ldw = cw.N;
    rowgemm_f32(p, c.s);
  }
}

void run_up(Ctx& c, const Stage& st, const void* x, int Lin, float* out, const float* bias) {
  const long N = (long)st.u * st.C;
  RowGemm p;
  p.x = x; p.x_bstride = (long)Lin * st.Cin; p.ldx = st.Cin; p.Lin = Lin;
  p.Cin = st.Cin; p.N = (int)N; p.taps = st.up.taps; p.dil = 1; p.B = c.B;
  p.out = out; p.o_bstride = (long)Lin * N; p.ldo = (int)N;
  if (st.up.taps == 2) {            // k = 2u: input row t feeds output rows t-1 (upper half of the kernel) and t
    p.center = 1; p.M = Lin + 1;
    p.o_shift = -(long)(st.u / 2) * st.C; p.o_limit = (long)Lin * N;
  } else {                          // k = u: one input sample -> u output samples, no overlap
    p.center = 0; p.M = Lin;
  }
  p.bias = bias;
  p.f16 = c.f16;
  ProfScope ps(c.e.prof, "bigvgan.ups", c.s);
  if (c.fast) {
    rowgemm_tc(p, st.up.tc[c.f16], c.s);
  } else {
    p.w = st.up.w.p; p.ldw = (int)N;
    rowgemm_f32(p, c.s);
  }
}

}  // namespace

void bigvgan_forward(Engine& e, BigVGANModel& m, const float* d_in, int B, int T, int precision, int16_t* d_pcm, float* d_wave,
                     const float* const* d_conds, long cl_bstride) {
  B2_CHECK(B > 0 && T > 0, "bigvgan: empty input");
  B2_CHECK(precision == PREC_F32 || precision == PREC_BF16 || precision == PREC_F16, "bigvgan: unknown precision");
  const bool latent_in = m.fn_w.p != nullptr;           // IndexTTS_F: d_in = GPT latent rows (T, gpt_dim), already channels-last
  B2_CHECK(!latent_in || B == 1, "the IndexTTS vocoder takes one latent sequence per call");
  B2_CHECK(latent_in == (d_conds != nullptr), "conditioning vectors go with the IndexTTS vocoder only");
  B2_CHECK(!latent_in || cl_bstride == 0, "the channels-last mel input belongs to the mel vocoder");
  cudaStream_t s = e.stream;
  Ctx c{e, m, B, precision != PREC_F32, precision == PREC_F16 ? 1 : 0, precision == PREC_F16 ? 2 : 1, s};
  // concurrent resblock branches: tensor-core engines only (the fp32 parity engine stays strictly serial)
  const bool par = c.fast && e.bigvgan_branches;
  if (par) e.ensure_aux();

  // workspace: every stage tensor has C*L <= C0*hop/… ; the largest is max_i(C_i * L_i), the post tensor adds 30 rows
  long maxel = (long)m.C0 * T;
  {
    long L = T;
    for (auto& st : m.stages) { L *= st.u; maxel = std::max(maxel, (long)st.C * L); }
    maxel = std::max(maxel, (long)m.Clast * (L + 30));
  }
  const size_t ws = (size_t)maxel * B;
  m.mel_cl.reserve((size_t)B * T * m.n_mels);
  m.xu.reserve(ws); m.xs.reserve(ws); m.abuf.reserve(ws);
  for (int j = 0; j < (par ? 3 : 1); ++j) { m.xa[j].reserve(ws); m.xb[j].reserve(ws); }
  if (c.fast) {
    for (int j = 0; j < (par ? 3 : 1); ++j) { m.abuf16[j].reserve(ws); m.cbuf16[j].reserve(ws); }
    m.mel16.reserve((size_t)B * T * round_up(m.n_mels, 8)); m.xs16.reserve(ws);
    bigvgan_tc_prepare(e, m, c.f16);  // 16-bit weight layouts (first fast call of each operand type only)
  } else {
    m.cbuf.reserve(ws);
  }

  const float* pre_bias = m.pre.bias.p;
  if (latent_in) {
    // latent = gpt.final_norm(hidden[:-2]) (Export_IndexTTS.py:301): T rows of gpt_dim, channels-last as the GEMM wants it;
    // conv_pre(.) + cond_layer vector, ups[i](.) + cond_i (:302,:306-307): the vectors are folded into the GEMM biases
    ProfScope ps(e.prof, "bigvgan.cond", s);
    layernorm_affine(d_in, m.fn_w.p, m.fn_b.p, m.mel_cl.p, T, m.n_mels, 1e-5f, s);
    cond_bias_kernel<<<ceil_div(m.C0, 128), 128, 0, s>>>(m.pre.bias.p, d_conds[m.nstages], m.pre_bias_c.p, m.C0, 1);
    B2_LAUNCH_CHECK(); count_launch();
    for (int i = 0; i < m.nstages; ++i) {
      Stage& st = m.stages[i];
      cond_bias_kernel<<<ceil_div(st.u * st.C, 128), 128, 0, s>>>(st.up.bias.p, d_conds[i], st.up_bias_c.p, st.C, st.u);
      B2_LAUNCH_CHECK(); count_launch();
    }
    pre_bias = m.pre_bias_c.p;
  } else if (cl_bstride > 0) {
    B2_CHECK(cl_bstride >= (long)T * m.n_mels, "bigvgan: channels-last batch stride shorter than one mel");
    ProfScope ps(e.prof, "bigvgan.mel_transpose", s);
    B2_CUDA(cudaMemcpy2DAsync(m.mel_cl.p, (size_t)T * m.n_mels * sizeof(float), d_in, (size_t)cl_bstride * sizeof(float),
                              (size_t)T * m.n_mels * sizeof(float), (size_t)B, cudaMemcpyDeviceToDevice, s));
  } else {
    ProfScope ps(e.prof, "bigvgan.mel_transpose", s);
    batched_transpose(d_in, m.mel_cl.p, B, m.n_mels, T, s);
  }
  const bool precise = !c.fast;
  const void* conv_in = m.mel_cl.p;
  int mel_ld = m.n_mels;
  if (c.fast) {
    ProfScope ps(e.prof, "bigvgan.cast", s);
    mel_ld = (int)round_up(m.n_mels, 8);      // bf16 rows must be 16-byte multiples for TMA
    cast_pad_f32_to_bf16(m.mel_cl.p, m.mel16.p, (long)B * T, m.n_mels, mel_ld, s, c.f16);
    conv_in = m.mel16.p;
  }
  // conv_pre -> xs (B, T, C0)
  // (the stage output is also written as bf16 by the producing GEMM: the next upsampler's A operand, no cast pass)
  run_conv(c, "bigvgan.conv_pre", m.pre, conv_in, T, 1, m.xs.p, 0, nullptr, 0, 1.0f, mel_ld, pre_bias, m.xs16.p);

  int L = T;
  static const char* kConvTag[8] = {"bigvgan.resconv.s0", "bigvgan.resconv.s1", "bigvgan.resconv.s2", "bigvgan.resconv.s3",
                                    "bigvgan.resconv.s4", "bigvgan.resconv.s5", "bigvgan.resconv.s6", "bigvgan.resconv.s7"};
  static const char* kActTag[8] = {"bigvgan.aa_snake.s0", "bigvgan.aa_snake.s1", "bigvgan.aa_snake.s2", "bigvgan.aa_snake.s3",
                                   "bigvgan.aa_snake.s4", "bigvgan.aa_snake.s5", "bigvgan.aa_snake.s6", "bigvgan.aa_snake.s7"};
  for (int i = 0; i < m.nstages; ++i) {
    Stage& st = m.stages[i];
    const char* ctag = kConvTag[i < 8 ? i : 7];
    const char* atag = kActTag[i < 8 ? i : 7];
    const void* up_in = c.fast ? (const void*)m.xs16.p : (const void*)m.xs.p;
    run_up(c, st, up_in, L, m.xu.p, latent_in ? st.up_bias_c.p : st.up.bias.p);
    L *= st.u;
    // fork: the three resblocks read the same stage input xu. Branch j enqueues on its own stream and owns its scratch tensors;
    // the branches meet again in xs, which they update in the reference's order ((b0 + b1) + b2) / 3: the last convolution of
    // branch j waits for that of branch j - 1 (an event), so the sum is bit-identical to the serial schedule.
    if (par) {
      B2_CUDA(cudaEventRecord(e.ev_fork, s));
      for (int a = 0; a < 2; ++a) B2_CUDA(cudaStreamWaitEvent(e.aux_stream[a], e.ev_fork, 0));
    }
    for (int j = 0; j < 3; ++j) {
      Ctx cj = c;
      cj.s = par && j > 0 ? e.aux_stream[j - 1] : s;
      const int bj = par ? j : 0;
      __nv_bfloat16* a16 = m.abuf16[bj].p;
      __nv_bfloat16* c16 = m.cbuf16[bj].p;
      const float* xcur = m.xu.p;
      for (int mm = 0; mm < 3; ++mm) {
        const bool last = (mm == 2);
        {
          ProfScope ps(e.prof, atag, cj.s);
          if (c.fast) aa_snake(xcur, 0, a16, c.t16, st.act[j][2 * mm].alpha.p, st.act[j][2 * mm].inv_beta.p, B, st.C, L, false, false, cj.s);
          else aa_snake(xcur, 0, m.abuf.p, 0, st.act[j][2 * mm].alpha.p, st.act[j][2 * mm].inv_beta.p, B, st.C, L, true, false, cj.s);
        }
        if (c.fast) run_conv(cj, ctag, st.c1[j][mm], a16, L, st.dil[j][mm], c16, c.t16, nullptr, 0, 1.0f);
        else run_conv(cj, ctag, st.c1[j][mm], m.abuf.p, L, st.dil[j][mm], m.cbuf.p, 0, nullptr, 0, 1.0f);
        {
          ProfScope ps(e.prof, atag, cj.s);
          if (c.fast) aa_snake(c16, c.t16, a16, c.t16, st.act[j][2 * mm + 1].alpha.p, st.act[j][2 * mm + 1].inv_beta.p, B, st.C, L, false, false, cj.s);
          else aa_snake(m.cbuf.p, 0, m.abuf.p, 0, st.act[j][2 * mm + 1].alpha.p, st.act[j][2 * mm + 1].inv_beta.p, B, st.C, L, true, false, cj.s);
        }
        const void* a2 = c.fast ? (const void*)a16 : (const void*)m.abuf.p;
        if (!last) {
          float* xnext = (mm == 0) ? m.xa[bj].p : m.xb[bj].p;
          run_conv(cj, ctag, st.c2[j][mm], a2, L, 1, xnext, 0, xcur, 0, 1.0f);   // x = xt + x
          xcur = xnext;
        } else {
          // xs (+)= conv + x ; the MRF mean (x 1/3, bigvgan.py:399) is folded into the third block's epilogue
          if (par && j > 0) B2_CUDA(cudaStreamWaitEvent(cj.s, e.ev_acc[j - 1], 0));
          run_conv(cj, ctag, st.c2[j][mm], a2, L, 1, m.xs.p, 0, xcur, j > 0 ? 1 : 0, j == 2 ? (float)(1.0 / 3.0) : 1.0f, 0, nullptr,
                   j == 2 ? m.xs16.p : nullptr);
          if (par && j < 2) B2_CUDA(cudaEventRecord(e.ev_acc[j], cj.s));
        }
      }
    }
    if (par) {      // join: the next stage (and the post-processing) continues on the caller's stream
      for (int a = 0; a < 2; ++a) {
        B2_CUDA(cudaEventRecord(e.ev_end[a], e.aux_stream[a]));
        B2_CUDA(cudaStreamWaitEvent(s, e.ev_end[a], 0));
      }
    }
  }
  {
    ProfScope ps(e.prof, "bigvgan.aa_snake_post", s);
    aa_snake(m.xs.p, 0, m.abuf.p, 0, m.post_act.alpha.p, m.post_act.inv_beta.p, B, m.Clast, L, precise, true, s);
  }
  {
    ProfScope ps(e.prof, "bigvgan.post_conv", s);
    const int Lo = L + 30;
    dim3 grid(ceil_div(Lo, 256), B);
    B2_CHECK(m.Clast == 24, "post_conv kernel is instantiated for 24 channels");
    post_conv_kernel<24><<<grid, 256, 0, s>>>(m.abuf.p, m.post_w.p, m.post_bias, Lo, d_pcm, d_wave);
    B2_LAUNCH_CHECK(); count_launch();
  }
}

}  // namespace b200tts

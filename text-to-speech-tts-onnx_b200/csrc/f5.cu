// F5-TTS on the GPU (channels-last / row-major [time][channel] everywhere).
//
// Reference semantics, quirks included:
//   graph A  F5_TTS/Export_F5.py:117-141  + dit.py:49-73 (TextEmbedding) + STFT_Process.py:144-157
//   graph B  F5_TTS/Export_F5.py:177-182  + dit.py:205-220 + modules.py:167-190,292-340,421-468,599-613
//   graph C  F5_TTS/Export_F5.py:197-203  + vocos/models.py:78-83, modules.py:43-51, heads.py:55-59 + STFT_Process.py:160-166
//
// What is restructured for the GPU (results unchanged):
//   * AdaLN: emb = Linear(SiLU(t)) acts on one of 32 constant rows -> all (step, layer) modulation vectors are
//     computed once at build time (mod tables); a step only runs the LayerNorm-modulate pass.
//   * Input embedding: W [x | c] = Wx x + Wc c; the Wc c + b half is step-invariant and computed once per utterance.
//   * q, k, v projections are one GEMM (N = 3072) whose epilogue applies the interleaved RoPE and writes V transposed.
//   * x += gate * y lives in the out-projection / FF2 epilogues; CFG + Euler is one element-wise kernel.
#include "f5.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "attention_tc.cuh"
#include "dit_chain.cuh"
#include "f5_kernels.cuh"
#include "layout.cuh"
#include "rowgemm.cuh"
#include "rowgemm_tc.cuh"

namespace b200tts {

namespace {

struct Lin {                 // y = x W^T + b ; reference weight [N][K]
  int K = 0, N = 0, Np = 0;  // Np = N rounded up to 4 (SIMT layout row stride)
  const float* w_ref = nullptr;
  DevBuf<float> w_own;       // when the reference-layout weight is assembled here (qkv concat, Wx/Wc split)
  DevBuf<float> wT;          // [K][Np] (lazy, fp32 engine)
  DevBuf<float> bias;        // [Np] (zero padded) or empty
  TcWeight tc[2];            // lazy, tensor-core engine: [0] bf16, [1] fp16 operands
};

struct GConv {               // grouped Conv1d (conv position embedding)
  int C = 0, k = 0, groups = 0;
  const float* w_ref = nullptr;   // (C, C/groups, k)
  DevBuf<float> w32;         // [g][j][c][n]
  DevBuf<float> bias;
  TcWeight tc[2];            // [g][j][n][c]: [0] bf16, [1] fp16
};

struct TextBlock {
  DevBuf<float> dw, dwb, lnw, lnb, gamma, beta;
  Lin pw1, pw2;
};
struct VocosBlock {
  DevBuf<float> dw, dwb, nw, nb;
  Lin pw1, pw2;
};
struct DiTLayer {
  Lin qkv, out, ff1, ff2;
  DevBuf<float> mod;         // [nfe][6*D]: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
  // LayerNorm folded into the GEMMs of the fused chain (dit_chain.cu), per 16-bit operand type ([0] bf16, [1] fp16):
  //   fu1 / fv1 [nfe][FF]: W_ff1 (1 + scale_mlp), W_ff1 shift_mlp + b_ff1;  fuq / fvq [nfe][3D]: the same for q|k|v with the msa pair
  DevBuf<float> fu1[2], fv1[2], fuq[2], fvq[2];
  bool fold_ready[2] = {false, false};
  // optional e4m3 mode of ff1 / q|k|v (engine option dit_fp8): weights quantised per output channel, their scales, and the
  // folded-LayerNorm vectors recomputed from the dequantised weights
  DevBuf<uint8_t> w8_ff1, w8_qkv, w8_ff2;
  DevBuf<float> sw_ff1, sw_qkv, sw_ff2, fu1_8, fv1_8, fuq_8, fvq_8, gate8, bias8;      // gate8 / bias8: ff2's gate_mlp / bias with the weight scale
  bool fp8_ready = false;
};

}  // namespace

struct F5Model {
  // hyper-parameters (from tensor shapes)
  int D = 0, H = 0, hd = 64, depth = 0, FF = 0, n_mels = 0, text_dim = 0, nfe = 0, nfft = 0, hop = 0, bins = 0;
  int cond_dim = 0;          // n_mels + text_dim
  int max_frames = 0;
  float cfg_strength = 2.0f;
  // DiT
  Lin wx, wc;
  GConv cp1, cp2;
  std::vector<DiTLayer> layers;
  DevBuf<float> mod_final;   // [nfe][2*D]: scale, shift
  Lin proj;
  std::vector<float> delta_t;
  const float *rope_cos = nullptr, *rope_sin = nullptr, *text_pos = nullptr, *text_table = nullptr;
  std::vector<TextBlock> text_blocks;
  // front end
  Lin stft;                  // K = nfft, N = 2*bins
  Lin fbank;                 // K = bins(+pad), N = n_mels
  // vocos + istft
  int VC = 0, VI = 0;
  DevBuf<float> v_embed_w, v_embed_b;       // [7][n_mels][VC]
  const float* v_embed_ref = nullptr;       // reference layout (VC, n_mels, 7): source of the tensor-core copy
  TcWeight v_embed_tc[2], istft_tc[2], head_tc[2];   // lazy, tensor-core engine ([0] bf16, [1] fp16)
  const float* istft_ref = nullptr;         // (2*bins, nfft)
  DevBuf<__nv_bfloat16> s16a, s16b;         // bf16 staging of the front / back end GEMM operands (tensor-core engine)
  DevBuf<float> v_nw, v_nb, v_fw, v_fb;
  std::vector<VocosBlock> vblocks;
  Lin head;                  // K = VC, N = 2*bins
  Lin istft;                 // K = 2*bins (+pad), N = nfft
  const float* wsi = nullptr;
  long wsi_len = 0;

  // ---- per-utterance state + workspaces (grow-only) ----
  int N = 0, ref_len = 0, Npad = 0;   // N = longest max_duration of the batch (every utterance's, when they are equal)
  int U = 1;                 // utterances in flight: rows of every DiT tensor are [u][cfg row][t], utterances concatenated
  // ragged batches (length-bucketed micro-batches; utterance u has Nu[u] frames, Fu[u] of them reference):
  std::vector<int> Nu, Fu;
  std::vector<long> tok_off; // rows of noise / cond before utterance u (sequence 2u+b starts at row 2*tok_off[u] + b*Nu[u])
  long Ntot = 0;             // sum of Nu
  bool ragged = false;       // the Nu differ: RoPE / V^T / attention take per-row (sequence, position) tables
  DevBuf<int> seq_off, seq_len;      // [2U] first row / length of every sequence
  DevBuf<int2> rowinfo;              // [2*Ntot] (sequence, position) of every DiT row
  std::vector<int> tables_for;       // the Nu the device tables were built for (skip the upload when a batch shape repeats)
  DevBuf<float> noise, cond, cond_drop, cproj, x, h, pred, rope_c, rope_s;
  const float *cur_cos = nullptr, *cur_sin = nullptr;
  DevBuf<float> n32, qkv32, att32, ff32, kT32, v32, s32, c32;          // fp32 engine
  DevBuf<__nv_bfloat16> h16, c16, n16, n16b, qk16, vT16, att16, ff16, x16;   // tensor-core engine (bf16 or fp16 bits)
  DevBuf<float> chain_stats;
  DevBuf<float> rowscale;       // [rows] 1 / std of every row at the latest LayerNorm (dit_chain.cu)                                           // dit_chain.cuh: team scratch
  DevBuf<unsigned> chain_flags;                                        // [depth][row blocks][8], zeroed once per Euler step
  DevBuf<unsigned long long> chain_trace;                              // debug only (B200TTS_CHAIN_TRACE)
  DevBuf<__half2> rope_cs16;                                           // [N][64] (cos, sin): exact, the tables are fp16-rounded (q5)
  // preprocess / decode scratch
  DevBuf<float> audio_f, spec, mag, mel, t_a, t_b, t_c, t_wide, grn_scratch;
  DevBuf<int> ids;
  DevBuf<float> d_a, d_b, d_c, d_wide, d_head, d_in, d_frames;
  bool tc_ready = false, f32_ready = false;
};

namespace {

void copy_vec(Engine& e, const std::string& name, DevBuf<float>& dst, long expect, long pad_to = 0) {
  const Tensor& t = e.weight(name);
  B2_CHECK(t.numel() == expect, name + ": unexpected size");
  const long n = pad_to > expect ? pad_to : expect;
  dst.alloc((size_t)n);
  B2_CUDA(cudaMemsetAsync(dst.p, 0, n * sizeof(float), e.stream));
  B2_CUDA(cudaMemcpyAsync(dst.p, t.data.p, expect * sizeof(float), cudaMemcpyDeviceToDevice, e.stream));
}

void lin_from(Engine& e, Lin& L, const std::string& wname, const std::string& bname) {
  const Tensor& W = e.weight(wname);
  B2_CHECK(W.shape.size() == 2, wname + " must be rank 2");
  L.N = (int)W.shape[0]; L.K = (int)W.shape[1]; L.Np = (int)round_up(L.N, 4);
  L.w_ref = W.data.p;
  if (!bname.empty()) copy_vec(e, bname, L.bias, L.N, L.Np);
}

void lin_prepare_f32(Engine& e, Lin& L) {
  if (L.wT.p) return;
  L.wT.alloc((size_t)L.K * L.Np);
  transpose_pad(L.w_ref, L.wT.p, L.N, L.K, L.Np, e.stream);
}
void lin_prepare_tc(Engine& e, Lin& L, int f16) {
  if (L.tc[f16].ready) return;
  tc_weight_from_f32(L.tc[f16], L.w_ref, 1, 1, L.N, L.K, e.stream, f16);
}

void dw_from(Engine& e, const std::string& wname, const std::string& bname, DevBuf<float>& w, DevBuf<float>& b, int C) {
  const Tensor& W = e.weight(wname);     // (C, 1, 7) -> [7][C]
  B2_CHECK(W.shape.size() == 3 && W.shape[0] == C && W.shape[1] == 1 && W.shape[2] == 7, wname + ": depthwise shape");
  w.alloc((size_t)7 * C);
  transpose_pad(W.data.p, w.p, C, 7, C, e.stream);
  copy_vec(e, bname, b, C);
}

struct Epi {
  const float* bias = nullptr; const float* gate = nullptr; const float* res = nullptr;
  int act = ACT_NONE; int out_bf16 = 0;
};

// fp32 [rows][C] (row stride ld) -> bf16 [rows][round_up(C, 8)] in `dst`: the A operand of a tensor-core GEMM whose producer
// is an fp32 kernel (norms, GRN, ISTFT input) -- the front / back end of the bf16 engine
const __nv_bfloat16* stage_bf16(Engine& e, DevBuf<__nv_bfloat16>& dst, const float* x, int ld, int rows, int C, int* ld16, int f16) {
  B2_CHECK(ld == C, "stage_bf16: compact rows expected");
  *ld16 = (int)round_up(C, 8);
  dst.reserve((size_t)rows * *ld16);
  ProfScope ps(e.prof, "f5.cast", e.stream);
  if (*ld16 == C) cast_f32_to_bf16(x, dst.p, (long)rows * C, e.stream, f16);
  else cast_pad_f32_to_bf16(x, dst.p, rows, C, *ld16, e.stream, f16);
  return dst.p;
}

// rows x K (ldx) @ W^T -> rows x N (ldo). use_tc: 0 = fp32 SIMT; 1 / 2 = bf16 / fp16 A operand + tcgen05 (engine.cuh Precision).
void linear(Engine& e, const char* tag, Lin& L, int use_tc, const void* x, int ldx, int rows, void* out, int ldo, const Epi& ep) {
  RowGemm p;
  p.x = x; p.x_bstride = 0; p.ldx = ldx; p.Lin = rows;
  p.Cin = L.K; p.N = use_tc ? L.N : L.Np; p.taps = 1; p.M = rows; p.B = 1;
  p.out = out; p.ldo = ldo; p.out_bf16 = ep.out_bf16;
  p.bias = ep.bias; p.gate = ep.gate; p.res = ep.res; p.act = ep.act;
  ProfScope ps(e.prof, tag, e.stream);
  if (use_tc) {
    const int f16 = use_tc == PREC_F16 ? 1 : 0;
    p.f16 = f16;
    lin_prepare_tc(e, L, f16);
    rowgemm_tc(p, L.tc[f16], e.stream);
  } else {
    lin_prepare_f32(e, L);
    B2_CHECK(L.Np <= ldo || L.Np == L.N, "linear: padded N exceeds the output row stride");
    p.w = L.wT.p; p.ldw = L.Np;
    rowgemm_f32(p, e.stream);
  }
}

}  // namespace

// =============================================================================================
// build
// =============================================================================================
F5Model* f5_build(Engine& e) {
  std::unique_ptr<F5Model> mp(new F5Model());
  F5Model& m = *mp;
  cudaStream_t s = e.stream;
  const std::string P = "dit.", V = "vocos.", C = "f5.";

  // ---- constants computed by the host exactly as the export script does (weights.py) ----
  {
    const Tensor& te = e.weight(C + "time_expand");      // (nfe, D)
    m.nfe = (int)te.shape[0]; m.D = (int)te.shape[1];
    const Tensor& dt = e.weight(C + "delta_t");
    B2_CHECK(dt.numel() == m.nfe - 1, "f5.delta_t must hold nfe-1 values");
    m.delta_t.resize(m.nfe - 1);
    B2_CUDA(cudaMemcpy(m.delta_t.data(), dt.data.p, (m.nfe - 1) * sizeof(float), cudaMemcpyDeviceToHost));
    const Tensor& rc = e.weight(C + "rope_cos");         // (max_frames, 64)
    m.max_frames = (int)rc.shape[0]; m.hd = (int)rc.shape[1];
    B2_CHECK(m.hd == 64, "head_dim must be 64");
    m.rope_cos = rc.data.p; m.rope_sin = e.weight(C + "rope_sin").data.p;
    m.text_pos = e.weight(C + "text_pos").data.p;
    m.H = m.D / m.hd;
  }
  // ---- DiT ----
  {
    const Tensor& W = e.weight(P + "input_embed.proj.weight");   // (D, 2*n_mels + text_dim)
    const Tensor& pw = e.weight(P + "proj_out.weight");          // (n_mels, D)
    m.n_mels = (int)pw.shape[0];
    const int kin = (int)W.shape[1];
    m.text_dim = kin - 2 * m.n_mels;
    m.cond_dim = m.n_mels + m.text_dim;
    B2_CHECK(W.shape[0] == m.D && m.text_dim > 0, "input_embed.proj.weight shape");
    B2_CHECK(m.n_mels % 4 == 0 && m.cond_dim % 4 == 0, "n_mels and n_mels+text_dim must be multiples of 4");
    // split W = [Wx | Wc] (columns): x part (n_mels), cond part (cond_dim)
    m.wx.N = m.D; m.wx.K = m.n_mels; m.wx.Np = m.D;
    m.wc.N = m.D; m.wc.K = m.cond_dim; m.wc.Np = m.D;
    m.wx.w_own.alloc((size_t)m.D * m.n_mels);
    m.wc.w_own.alloc((size_t)m.D * m.cond_dim);
    B2_CUDA(cudaMemcpy2DAsync(m.wx.w_own.p, m.n_mels * sizeof(float), W.data.p, kin * sizeof(float), m.n_mels * sizeof(float), m.D, cudaMemcpyDeviceToDevice, s));
    B2_CUDA(cudaMemcpy2DAsync(m.wc.w_own.p, m.cond_dim * sizeof(float), W.data.p + m.n_mels, kin * sizeof(float), m.cond_dim * sizeof(float), m.D, cudaMemcpyDeviceToDevice, s));
    m.wx.w_ref = m.wx.w_own.p; m.wc.w_ref = m.wc.w_own.p;
    copy_vec(e, P + "input_embed.proj.bias", m.wc.bias, m.D);
  }
  for (int which = 0; which < 2; ++which) {
    GConv& g = which == 0 ? m.cp1 : m.cp2;
    const std::string n = P + "input_embed.conv_pos_embed.conv1d." + std::to_string(which * 2);
    const Tensor& W = e.weight(n + ".weight");           // (D, D/groups, k)
    g.C = (int)W.shape[0]; g.k = (int)W.shape[2]; g.groups = g.C / (int)W.shape[1];
    B2_CHECK(g.C == m.D && g.k % 2 == 1, "conv_pos_embed weight shape");
    g.w_ref = W.data.p;
    copy_vec(e, n + ".bias", g.bias, m.D);
  }
  while (e.has_weight(P + "transformer_blocks." + std::to_string(m.depth) + ".attn.to_q.weight")) ++m.depth;
  B2_CHECK(m.depth > 0, "no dit.transformer_blocks.* tensors loaded");
  m.layers.resize(m.depth);
  // SiLU(time_expand) once, then every AdaLN linear on all nfe rows
  DevBuf<float> st((size_t)m.nfe * m.D);
  silu(e.weight(C + "time_expand").data.p, st.p, (long)m.nfe * m.D, s);
  auto adaln = [&](const std::string& name, int out_dim, DevBuf<float>& table) {
    Lin L;
    lin_from(e, L, name + ".weight", name + ".bias");
    B2_CHECK(L.N == out_dim && L.K == m.D, name + ": AdaLN linear shape");
    table.alloc((size_t)m.nfe * out_dim);
    Epi ep; ep.bias = L.bias.p;
    linear(e, "f5.build", L, 0, st.p, m.D, m.nfe, table.p, out_dim, ep);
    B2_CUDA(cudaStreamSynchronize(s));       // L (and its transposed copy) is freed at scope exit
  };
  for (int i = 0; i < m.depth; ++i) {
    DiTLayer& L = m.layers[i];
    const std::string bp = P + "transformer_blocks." + std::to_string(i) + ".";
    adaln(bp + "attn_norm.linear", 6 * m.D, L.mod);
    // fused q|k|v weight (3D, D) and bias
    L.qkv.N = 3 * m.D; L.qkv.K = m.D; L.qkv.Np = 3 * m.D;
    L.qkv.w_own.alloc((size_t)3 * m.D * m.D);
    L.qkv.bias.alloc((size_t)3 * m.D);
    const char* names[3] = {"to_q", "to_k", "to_v"};
    for (int j = 0; j < 3; ++j) {
      const Tensor& W = e.weight(bp + "attn." + names[j] + ".weight");
      const Tensor& b = e.weight(bp + "attn." + names[j] + ".bias");
      B2_CHECK(W.numel() == (long)m.D * m.D && b.numel() == m.D, "attention projection shape");
      B2_CUDA(cudaMemcpyAsync(L.qkv.w_own.p + (size_t)j * m.D * m.D, W.data.p, (size_t)m.D * m.D * sizeof(float), cudaMemcpyDeviceToDevice, s));
      B2_CUDA(cudaMemcpyAsync(L.qkv.bias.p + (size_t)j * m.D, b.data.p, m.D * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    L.qkv.w_ref = L.qkv.w_own.p;
    lin_from(e, L.out, bp + "attn.to_out.0.weight", bp + "attn.to_out.0.bias");
    lin_from(e, L.ff1, bp + "ff.ff.0.0.weight", bp + "ff.ff.0.0.bias");
    lin_from(e, L.ff2, bp + "ff.ff.2.weight", bp + "ff.ff.2.bias");
    m.FF = L.ff1.N;
  }
  adaln(P + "norm_out.linear", 2 * m.D, m.mod_final);
  lin_from(e, m.proj, P + "proj_out.weight", P + "proj_out.bias");
  // ---- text embedding ----
  {
    const Tensor& T = e.weight(P + "text_embed.text_embed.weight");
    B2_CHECK(T.shape.size() == 2 && T.shape[1] == m.text_dim, "text embedding table shape");
    m.text_table = T.data.p;
    int nb = 0;
    while (e.has_weight(P + "text_embed.text_blocks." + std::to_string(nb) + ".dwconv.weight")) ++nb;
    m.text_blocks.resize(nb);
    for (int i = 0; i < nb; ++i) {
      TextBlock& tb = m.text_blocks[i];
      const std::string bp = P + "text_embed.text_blocks." + std::to_string(i) + ".";
      dw_from(e, bp + "dwconv.weight", bp + "dwconv.bias", tb.dw, tb.dwb, m.text_dim);
      copy_vec(e, bp + "norm.weight", tb.lnw, m.text_dim);
      copy_vec(e, bp + "norm.bias", tb.lnb, m.text_dim);
      lin_from(e, tb.pw1, bp + "pwconv1.weight", bp + "pwconv1.bias");
      lin_from(e, tb.pw2, bp + "pwconv2.weight", bp + "pwconv2.bias");
      copy_vec(e, bp + "grn.gamma", tb.gamma, tb.pw1.N);
      copy_vec(e, bp + "grn.beta", tb.beta, tb.pw1.N);
    }
  }
  // ---- STFT front end: basis [2*bins][nfft] = [cos ; sin] kernels, fbank (n_mels, bins) ----
  {
    const Tensor& B = e.weight(C + "stft_basis");        // (2*bins, nfft)
    m.nfft = (int)B.shape[1]; m.bins = (int)B.shape[0] / 2; m.hop = m.nfft / 4;
    m.stft.N = 2 * m.bins; m.stft.K = m.nfft; m.stft.Np = (int)round_up(2 * m.bins, 4); m.stft.w_ref = B.data.p;
    const Tensor& Fb = e.weight(C + "fbank");            // (n_mels, bins) -> pad K to a multiple of 4
    B2_CHECK(Fb.shape[0] == m.n_mels && Fb.shape[1] == m.bins, "f5.fbank shape");
    const int kp = (int)round_up(m.bins, 4);
    m.fbank.N = m.n_mels; m.fbank.K = kp; m.fbank.Np = m.n_mels;
    m.fbank.w_own.alloc((size_t)m.n_mels * kp);
    B2_CUDA(cudaMemsetAsync(m.fbank.w_own.p, 0, (size_t)m.n_mels * kp * sizeof(float), s));
    B2_CUDA(cudaMemcpy2DAsync(m.fbank.w_own.p, kp * sizeof(float), Fb.data.p, m.bins * sizeof(float), m.bins * sizeof(float), m.n_mels, cudaMemcpyDeviceToDevice, s));
    m.fbank.w_ref = m.fbank.w_own.p;
  }
  // ---- Vocos (weights already folded by the host as Export_F5.py:390-402 does) + ISTFT tables ----
  {
    const Tensor& W = e.weight(V + "backbone.embed.weight");     // (VC, n_mels, 7)
    m.VC = (int)W.shape[0];
    B2_CHECK(W.shape[1] == m.n_mels && W.shape[2] == 7, "vocos embed shape");
    m.v_embed_w.alloc((size_t)7 * m.n_mels * m.VC);
    conv_weight_permute(W.data.p, m.v_embed_w.p, m.VC, m.n_mels, 7, 1, 1, s);
    m.v_embed_ref = W.data.p;
    copy_vec(e, V + "backbone.embed.bias", m.v_embed_b, m.VC);
    copy_vec(e, V + "backbone.norm.weight", m.v_nw, m.VC);
    copy_vec(e, V + "backbone.norm.bias", m.v_nb, m.VC);
    copy_vec(e, V + "backbone.final_layer_norm.weight", m.v_fw, m.VC);
    copy_vec(e, V + "backbone.final_layer_norm.bias", m.v_fb, m.VC);
    int nb = 0;
    while (e.has_weight(V + "backbone.convnext." + std::to_string(nb) + ".dwconv.weight")) ++nb;
    m.vblocks.resize(nb);
    for (int i = 0; i < nb; ++i) {
      VocosBlock& vb = m.vblocks[i];
      const std::string bp = V + "backbone.convnext." + std::to_string(i) + ".";
      dw_from(e, bp + "dwconv.weight", bp + "dwconv.bias", vb.dw, vb.dwb, m.VC);
      copy_vec(e, bp + "norm.weight", vb.nw, m.VC);
      copy_vec(e, bp + "norm.bias", vb.nb, m.VC);
      lin_from(e, vb.pw1, bp + "pwconv1.weight", bp + "pwconv1.bias");
      lin_from(e, vb.pw2, bp + "pwconv2.weight", bp + "pwconv2.bias");
      B2_CHECK(!e.has_weight(bp + "gamma"), "vocos gamma must be folded into pwconv2 by the host (Export_F5.py:401-402)");
      m.VI = vb.pw1.N;
    }
    lin_from(e, m.head, V + "head.out.weight", V + "head.out.bias");
    B2_CHECK(m.head.N == 2 * m.bins, "vocos head must emit n_fft + 2 rows");
    const Tensor& IB = e.weight(C + "istft_basis");      // (2*bins, nfft): frame = inp (1 x 2*bins) @ IB
    B2_CHECK(IB.shape[0] == 2 * m.bins && IB.shape[1] == m.nfft, "f5.istft_basis shape");
    // as a Lin: y = x W^T with W = IB^T (nfft, 2*bins); the SIMT layout wT = [K][N] = IB itself, K padded to 4
    const int kp = (int)round_up(2 * m.bins, 4);
    m.istft.N = m.nfft; m.istft.K = kp; m.istft.Np = m.nfft;
    m.istft.wT.alloc((size_t)kp * m.nfft);
    B2_CUDA(cudaMemsetAsync(m.istft.wT.p, 0, (size_t)kp * m.nfft * sizeof(float), s));
    B2_CUDA(cudaMemcpyAsync(m.istft.wT.p, IB.data.p, (size_t)2 * m.bins * m.nfft * sizeof(float), cudaMemcpyDeviceToDevice, s));
    m.istft_ref = IB.data.p;
    const Tensor& ws = e.weight(C + "window_sum_inv");
    m.wsi = ws.data.p; m.wsi_len = ws.numel();
  }
  B2_CUDA(cudaStreamSynchronize(s));
  return mp.release();
}

void f5_free(F5Model* m) { delete m; }

// =============================================================================================
// state helpers
// =============================================================================================
static F5Model& model(Engine& e) {
  B2_CHECK(e.f5 != nullptr, "F5 weights are not built (call b200tts_f5_build)");
  return *e.f5;
}
int f5_ref_len(const Engine& e) { B2_CHECK(e.f5 != nullptr, "F5 not built"); return e.f5->ref_len; }
int f5_seq_len(const Engine& e) { B2_CHECK(e.f5 != nullptr, "F5 not built"); return e.f5->N; }
int f5_cond_dim(const Engine& e) { B2_CHECK(e.f5 != nullptr, "F5 not built"); return e.f5->cond_dim; }
int f5_n_mels(const Engine& e) { B2_CHECK(e.f5 != nullptr, "F5 not built"); return e.f5->n_mels; }
int f5_nfe(const Engine& e) { B2_CHECK(e.f5 != nullptr, "F5 not built"); return e.f5->nfe; }
static long tok_of(const F5Model& m, int u) {
  B2_CHECK(u >= 0 && u < (int)m.tok_off.size(), "utterance index outside the batch");
  return m.tok_off[u];
}
float* f5_cond(Engine& e, int u) { F5Model& m = model(e); return m.cond.p + (size_t)tok_of(m, u) * m.cond_dim; }
float* f5_cond_drop(Engine& e, int u) { F5Model& m = model(e); return m.cond_drop.p + (size_t)tok_of(m, u) * m.cond_dim; }
float* f5_noise(Engine& e, int u) { F5Model& m = model(e); return m.noise.p + (size_t)tok_of(m, u) * m.n_mels; }
long f5_tok_offset(const Engine& e, int u) { B2_CHECK(e.f5 != nullptr, "F5 not built"); return tok_of(*e.f5, u); }

// host-side shape state of a batch: per-utterance lengths, offsets, totals (no device work)
static void set_shape(F5Model& m, int U, const int* Ns, const int* Fs) {
  B2_CHECK(U >= 1 && U <= 4096, "batch of utterances must be in [1, 4096]");
  m.U = U; m.Nu.assign(Ns, Ns + U); m.Fu.assign(U, 0); m.tok_off.assign(U, 0);
  m.N = 0; m.Ntot = 0; m.ragged = false;
  for (int u = 0; u < U; ++u) {
    B2_CHECK(Ns[u] > 0 && Ns[u] <= m.max_frames, "max_duration must be in [1, MAX_SIGNAL_LENGTH]");
    if (Fs) m.Fu[u] = Fs[u];
    m.tok_off[u] = m.Ntot;
    m.Ntot += Ns[u];
    if (Ns[u] > m.N) m.N = Ns[u];
    if (Ns[u] != Ns[0]) m.ragged = true;
  }
  m.Npad = (int)round_up(m.N, 8);
  m.ref_len = m.Fu[0];
  m.cur_cos = m.rope_cos; m.cur_sin = m.rope_sin;     // rows [0, N) of the fp16-rounded tables
}

void f5_begin_ragged(Engine& e, int U, const int* Ns) {
  F5Model& m = model(e);
  set_shape(m, U, Ns, nullptr);
  const size_t T = (size_t)m.Ntot, R = 2 * T;
  m.noise.reserve(T * m.n_mels);
  m.cond.reserve(T * m.cond_dim);
  m.cond_drop.reserve(T * m.cond_dim);
  m.cproj.reserve(R * m.D);
  m.x.reserve(R * m.D);
  m.h.reserve(R * m.D);
  m.pred.reserve(R * m.n_mels);
  if (m.ragged && m.tables_for != m.Nu) {   // per-sequence / per-row tables (host -> device: call this OUTSIDE a stream capture)
    std::vector<int> so(2 * U), sl(2 * U);
    std::vector<int2> ri(R);
    for (int u = 0; u < U; ++u)
      for (int b = 0; b < 2; ++b) {
        const int sq = 2 * u + b;
        so[sq] = (int)(2 * m.tok_off[u] + (long)b * m.Nu[u]); sl[sq] = m.Nu[u];
        for (int t = 0; t < m.Nu[u]; ++t) ri[(size_t)so[sq] + t] = make_int2(sq, t);
      }
    m.seq_off.reserve(so.size()); m.seq_len.reserve(sl.size()); m.rowinfo.reserve(ri.size());
    B2_CUDA(cudaMemcpyAsync(m.seq_off.p, so.data(), so.size() * sizeof(int), cudaMemcpyHostToDevice, e.stream));
    B2_CUDA(cudaMemcpyAsync(m.seq_len.p, sl.data(), sl.size() * sizeof(int), cudaMemcpyHostToDevice, e.stream));
    B2_CUDA(cudaMemcpyAsync(m.rowinfo.p, ri.data(), ri.size() * sizeof(int2), cudaMemcpyHostToDevice, e.stream));
    B2_CUDA(cudaStreamSynchronize(e.stream));      // the host vectors go out of scope
    m.tables_for = m.Nu;
  }
}

void f5_begin(Engine& e, int N, int U) {
  std::vector<int> Ns((size_t)(U > 0 ? U : 1), N);
  f5_begin_ragged(e, U, Ns.data());
}

void f5_set_rope(Engine& e, const float* d_cos, const float* d_sin) {
  F5Model& m = model(e);
  m.rope_c.reserve((size_t)m.N * m.hd); m.rope_s.reserve((size_t)m.N * m.hd);
  B2_CUDA(cudaMemcpyAsync(m.rope_c.p, d_cos, (size_t)m.N * m.hd * sizeof(float), cudaMemcpyDeviceToDevice, e.stream));
  B2_CUDA(cudaMemcpyAsync(m.rope_s.p, d_sin, (size_t)m.N * m.hd * sizeof(float), cudaMemcpyDeviceToDevice, e.stream));
  m.cur_cos = m.rope_c.p; m.cur_sin = m.rope_s.p;
}

void f5_rope_buffers(Engine& e, float** d_cos, float** d_sin) {
  F5Model& m = model(e);
  m.rope_c.reserve((size_t)m.N * m.hd); m.rope_s.reserve((size_t)m.N * m.hd);
  m.cur_cos = m.rope_c.p; m.cur_sin = m.rope_s.p;
  *d_cos = m.rope_c.p; *d_sin = m.rope_s.p;
}

void f5_restore_shape(Engine& e, int N, int ref_len, int U) {
  std::vector<int> Ns((size_t)U, N), Fs((size_t)U, ref_len);
  set_shape(model(e), U, Ns.data(), Fs.data());
}
void f5_restore_ragged(Engine& e, int U, const int* Ns, const int* Fs) { set_shape(model(e), U, Ns, Fs); }

void f5_prepare_cond(Engine& e) {
  F5Model& m = model(e);
  Epi ep; ep.bias = m.wc.bias.p;
  for (int u = 0; u < m.U; ++u) {             // sequence 2u = cond, 2u+1 = cond_drop of utterance u
    const int Nu = m.Nu[u];
    const size_t co = (size_t)m.tok_off[u] * m.cond_dim, po = (size_t)2 * m.tok_off[u] * m.D;
    linear(e, "f5.cond_proj", m.wc, 0, m.cond.p + co, m.cond_dim, Nu, m.cproj.p + po, m.D, ep);
    linear(e, "f5.cond_proj", m.wc, 0, m.cond_drop.p + co, m.cond_dim, Nu, m.cproj.p + po + (size_t)Nu * m.D, m.D, ep);
  }
}

// =============================================================================================
// graph A
// =============================================================================================
void f5_preprocess(Engine& e, const int16_t* d_audio, long L, const int* d_text_ids, int n_text, int N, int u, int U, int fast) {
  F5Model& m = model(e);
  cudaStream_t s = e.stream;
  const int F = (int)(L / m.hop) + 1;
  B2_CHECK(L > m.nfft / 2, "audio too short for the reflect-padded STFT");
  B2_CHECK(n_text >= 0 && n_text <= N, "text_ids longer than max_duration");
  B2_CHECK(F <= N, "max_duration shorter than the reference audio");
  B2_CHECK(u >= 0 && u < U, "utterance index outside the batch");
  if (U == 1) f5_begin(e, N, 1);              // a batch (U > 1) is laid out by f5_begin / f5_begin_ragged before its first utterance
  B2_CHECK(m.U == U && (int)m.Nu.size() == U && m.Nu[u] == N, "f5_preprocess: the batch layout (f5_begin_ragged) does not match this utterance");
  m.Fu[u] = F;
  if (u == 0) m.ref_len = F;
  float* cond = m.cond.p + (size_t)m.tok_off[u] * m.cond_dim;
  float* cond_drop = m.cond_drop.p + (size_t)m.tok_off[u] * m.cond_dim;
  // ---- STFT -> |X| -> mel -> log : cond[:, :n_mels] ----
  {
    ProfScope ps(e.prof, "f5.pre_elementwise", s);
    m.audio_f.reserve((size_t)L + m.nfft);
    audio_reflect_pad(d_audio, m.audio_f.p, L, m.nfft / 2, s);
  }
  m.spec.reserve((size_t)F * m.stft.Np);
  {
    // frames are overlapping rows of the padded signal: row stride = hop (STFT_Process.py:153-157 as one GEMM)
    Epi ep;
    linear(e, "f5.stft", m.stft, 0, m.audio_f.p, m.hop, F, m.spec.p, m.stft.Np, ep);
  }
  const int ldm = m.fbank.K;
  m.mag.reserve((size_t)F * ldm);
  m.mel.reserve((size_t)N * m.n_mels);
  {
    ProfScope ps(e.prof, "f5.pre_elementwise", s);
    stft_magnitude(m.spec.p, m.stft.Np, m.mag.p, ldm, F, m.bins, s);
  }
  {
    Epi ep;
    linear(e, "f5.fbank", m.fbank, 0, m.mag.p, ldm, F, m.mel.p, m.n_mels, ep);
  }
  {
    ProfScope ps(e.prof, "f5.pre_elementwise", s);
    logmel_into(m.mel.p, F, cond, m.cond_dim, 0, N, m.n_mels, s);
    copy_cols(nullptr, 0, cond_drop, m.cond_dim, 0, N, m.n_mels, s);         // zeros
  }
  // ---- text embedding (dit.py:49-73), text then text_drop ----
  const int TD = m.text_dim, TW = m.text_blocks.empty() ? TD : m.text_blocks[0].pw1.N;
  m.ids.reserve((size_t)N);
  m.t_a.reserve((size_t)N * TD); m.t_b.reserve((size_t)N * TD); m.t_c.reserve((size_t)N * TD);
  m.t_wide.reserve((size_t)N * TW); m.grn_scratch.reserve((size_t)TW + 8);
  {
    ProfScope ps(e.prof, "f5.pre_elementwise", s);
    pad_text_ids(d_text_ids, n_text, m.ids.p, N, s);
  }
  for (int drop = 0; drop < 2; ++drop) {
    ProfScope ps(e.prof, "f5.text_embed", s);
    text_embed_gather(m.ids.p, m.text_table, m.text_pos, m.t_a.p, N, TD, drop == 0 ? 1 : 0, s);
    for (auto& tb : m.text_blocks) {
      dwconv7(m.t_a.p, tb.dw.p, tb.dwb.p, m.t_b.p, 1, N, TD, s);
      layernorm_affine(m.t_b.p, tb.lnw.p, tb.lnb.p, m.t_c.p, N, TD, 1e-6f, s);
      Epi e1; e1.bias = tb.pw1.bias.p; e1.act = ACT_GELU_ERF;
      Epi e2; e2.bias = tb.pw2.bias.p; e2.res = m.t_a.p;
      if (fast) {                      // bf16 engine: the two pointwise GEMMs on tensor cores (GRN needs the fp32 hidden)
        int l16 = 0;
        const int f16 = fast == PREC_F16 ? 1 : 0;
        const __nv_bfloat16* a1 = stage_bf16(e, m.s16a, m.t_c.p, TD, N, TD, &l16, f16);
        linear(e, "f5.text_embed.gemm", tb.pw1, fast, a1, l16, N, m.t_wide.p, TW, e1);
        grn_inplace(m.t_wide.p, tb.gamma.p, tb.beta.p, m.grn_scratch.p, N, TW, s);
        const __nv_bfloat16* a2 = stage_bf16(e, m.s16b, m.t_wide.p, TW, N, TW, &l16, f16);
        linear(e, "f5.text_embed.gemm", tb.pw2, fast, a2, l16, N, m.t_b.p, TD, e2);
      } else {
        linear(e, "f5.text_embed.gemm", tb.pw1, 0, m.t_c.p, TD, N, m.t_wide.p, TW, e1);
        grn_inplace(m.t_wide.p, tb.gamma.p, tb.beta.p, m.grn_scratch.p, N, TW, s);
        linear(e, "f5.text_embed.gemm", tb.pw2, 0, m.t_wide.p, TW, N, m.t_b.p, TD, e2);
      }
      mask_rows(m.t_b.p, m.ids.p, N, TD, s);
      std::swap(m.t_a, m.t_b);
    }
    copy_cols(m.t_a.p, TD, drop == 0 ? cond : cond_drop, m.cond_dim, m.n_mels, N, TD, s);
  }
}

// =============================================================================================
// graph B
// =============================================================================================
namespace {

void gconv(Engine& e, GConv& g, int fast, const void* x, int N, int nseq, void* out, int out_bf16, int act, const float* res) {
  const int cg = g.C / g.groups;
  RowGemm p;
  p.x = x; p.x_bstride = (long)N * g.C; p.ldx = g.C; p.Lin = N;
  p.Cin = cg; p.N = cg; p.taps = g.k; p.dil = 1; p.center = (g.k - 1) / 2; p.groups = g.groups;
  p.M = N; p.B = nseq;
  p.out = out; p.o_bstride = (long)N * g.C; p.ldo = g.C; p.out_bf16 = out_bf16;
  p.bias = g.bias.p; p.act = act; p.res = res;
  ProfScope ps(e.prof, "f5.conv_pos", e.stream);
  if (fast) {
    const int f16 = fast == PREC_F16 ? 1 : 0;
    p.f16 = f16;
    if (!g.tc[f16].ready) {
      DevBuf<float> tmp((size_t)g.C * cg * g.k);
      conv_weight_permute(g.w_ref, tmp.p, g.C, cg, g.k, g.groups, 0, e.stream);
      tc_weight_from_f32(g.tc[f16], tmp.p, g.groups, g.k, cg, cg, e.stream, f16);
      B2_CUDA(cudaStreamSynchronize(e.stream));
    }
    rowgemm_tc(p, g.tc[f16], e.stream);
  } else {
    if (!g.w32.p) {
      g.w32.alloc((size_t)g.C * cg * g.k);
      conv_weight_permute(g.w_ref, g.w32.p, g.C, cg, g.k, g.groups, 1, e.stream);
    }
    p.w = g.w32.p; p.ldw = cg;
    rowgemm_f32(p, e.stream);
  }
}

void reserve_step(F5Model& m, bool fast) {
  const size_t R = (size_t)2 * m.Ntot;
  if (fast) {
    m.h16.reserve(R * m.D); m.c16.reserve(R * m.D); m.n16.reserve(R * m.D); m.n16b.reserve(R * m.D); m.att16.reserve(R * m.D);
    m.chain_stats.reserve(dit_chain_stats_floats((int)R));
    m.rowscale.reserve(R);
    m.chain_flags.reserve(dit_chain_flag_words((int)R) * (size_t)m.depth);
    m.qk16.reserve(R * 2 * m.D); m.vT16.reserve((size_t)2 * m.U * m.H * m.hd * m.Npad); m.ff16.reserve(R * m.FF);
    m.rope_cs16.reserve((size_t)m.N * m.hd);
    m.x16.reserve((m.ragged ? R : (size_t)m.Ntot) * round_up(m.n_mels, 8));    // ragged: one x row per DiT row (ragged_embed)
  } else {
    m.c32.reserve(R * m.D); m.n32.reserve(R * m.D); m.att32.reserve(R * m.D); m.qkv32.reserve(R * 3 * m.D);
    m.ff32.reserve(R * m.FF);
    m.kT32.reserve((size_t)2 * m.H * m.hd * m.Npad); m.v32.reserve((size_t)2 * m.H * m.Npad * m.hd);
    m.s32.reserve((size_t)m.N * m.H * m.Npad);
  }
}

// fp32 attention for one CFG row b via two grouped GEMMs and a row softmax (parity engine)
void attention_f32(Engine& e, F5Model& m, int b) {
  cudaStream_t s = e.stream;
  const int N = m.N, H = m.H, hd = m.hd, Np = m.Npad;
  ProfScope ps(e.prof, "f5.attention_f32", s);
  RowGemm p;                                    // S[t][h*Np + key] = q[t][h*64 + :] . kT[b][h][:][key]
  p.x = m.qkv32.p + (size_t)b * N * 3 * m.D; p.ldx = 3 * m.D; p.Lin = N;
  p.Cin = hd; p.N = Np; p.taps = 1; p.groups = H; p.M = N; p.B = 1;
  p.w = m.kT32.p + (size_t)b * H * hd * Np; p.ldw = Np;
  p.out = m.s32.p; p.ldo = H * Np;
  rowgemm_f32(p, s);
  softmax_rows(m.s32.p, (long)N * H, N, Np, s);
  RowGemm o;                                    // O[t][h*64 + d] = P[t][h*Np + :] . v[b][h][:][d]
  o.x = m.s32.p; o.ldx = H * Np; o.Lin = N;
  o.Cin = Np; o.N = hd; o.taps = 1; o.groups = H; o.M = N; o.B = 1;
  o.w = m.v32.p + (size_t)b * H * Np * hd; o.ldw = hd;      // v is [2][H][Npad][hd], padding rows zero
  o.out = m.att32.p + (size_t)b * N * m.D; o.ldo = m.D;
  rowgemm_f32(o, s);
}

}  // namespace

void f5_steps(Engine& e, int first, int count, int precision) {
  F5Model& m = model(e);
  cudaStream_t s = e.stream;
  B2_CHECK(m.N > 0, "f5_steps: no utterance state (run preprocess / begin first)");
  B2_CHECK(first >= 0 && count >= 0 && first + count <= m.nfe - 1, "f5_steps: time_step out of range");
  B2_CHECK(precision == PREC_F32 || precision == PREC_BF16 || precision == PREC_F16, "f5_steps: unknown precision");
  const int fast = precision == PREC_F32 ? 0 : precision;      // 0 = fp32 SIMT parity engine, 1 / 2 = bf16 / fp16 tensor-core engine
  const int f16 = precision == PREC_F16 ? 1 : 0;
  const int N = m.N, D = m.D, S = 2 * m.U, R = (int)(2 * m.Ntot);   // N = the longest utterance
  const int2* rowinfo = m.ragged ? m.rowinfo.p : nullptr;
  B2_CHECK(fast || m.U == 1, "the fp32 parity engine runs one utterance at a time");
  B2_CHECK((int)m.Nu.size() == m.U && m.Ntot > 0, "f5_steps: no batch layout (f5_begin / f5_begin_ragged)");
  reserve_step(m, fast != 0);
  // e.dit_chain = false (B200TTS_CHAIN=0 / b200tts_set_option) keeps every DiT block as seven launches (LN, q|k|v, attention, out,
  // LN, ff1, ff2): the A/B switch for the fused row-block chain (dit_chain.cu)
  const bool chain = fast && e.dit_chain && dit_chain_supported(D, m.FF, m.H);
  PdlScope pdl(m.U == 1);                  // short kernels only (common.cuh)
  if (!fast) {   // padding rows / columns (t in [N, Npad)) of the fp32 attention operands must read as zero
    B2_CUDA(cudaMemsetAsync(m.kT32.p, 0, (size_t)2 * m.H * m.hd * m.Npad * sizeof(float), s));
    B2_CUDA(cudaMemsetAsync(m.v32.p, 0, (size_t)2 * m.H * m.Npad * m.hd * sizeof(float), s));
  }
  if (fast) {
    ProfScope ps(e.prof, "f5.cast", s);
    // stale V^T columns beyond a short sequence's length meet probabilities of exactly zero: they must hold finite values
    if (m.ragged) B2_CUDA(cudaMemsetAsync(m.vT16.p, 0, (size_t)2 * m.U * m.H * m.hd * m.Npad * sizeof(__nv_bfloat16), s));
    rope_pack_half(m.cur_cos, m.cur_sin, m.rope_cs16.p, (long)N * m.hd, s);
    for (auto& L : m.layers) { lin_prepare_tc(e, L.qkv, f16); lin_prepare_tc(e, L.out, f16); lin_prepare_tc(e, L.ff1, f16); lin_prepare_tc(e, L.ff2, f16); }
    if (chain) {
      for (auto& L : m.layers) {
        if (L.fold_ready[f16]) continue;
        L.fu1[f16].alloc((size_t)m.nfe * m.FF); L.fv1[f16].alloc((size_t)m.nfe * m.FF);
        L.fuq[f16].alloc((size_t)m.nfe * 3 * D); L.fvq[f16].alloc((size_t)m.nfe * 3 * D);
        fold_vectors(L.ff1.tc[f16].w.p, L.ff1.tc[f16].ldc, D, f16, L.mod.p + 4 * D, L.mod.p + 3 * D, 6 * D, L.ff1.bias.p, L.fu1[f16].p, L.fv1[f16].p,
                     m.FF, m.nfe, s);
        fold_vectors(L.qkv.tc[f16].w.p, L.qkv.tc[f16].ldc, D, f16, L.mod.p + D, L.mod.p, 6 * D, L.qkv.bias.p, L.fuq[f16].p, L.fvq[f16].p, 3 * D,
                     m.nfe, s);
        L.fold_ready[f16] = true;
      }
      if (e.dit_fp8) {
        for (auto& L : m.layers) {
          if (L.fp8_ready) continue;
          L.w8_ff1.alloc((size_t)m.FF * D); L.sw_ff1.alloc(m.FF); L.w8_qkv.alloc((size_t)3 * D * D); L.sw_qkv.alloc((size_t)3 * D);
          quantize_rows_e4m3(L.ff1.w_ref, m.FF, D, L.w8_ff1.p, D, L.sw_ff1.p, s);
          quantize_rows_e4m3(L.qkv.w_ref, 3 * D, D, L.w8_qkv.p, D, L.sw_qkv.p, s);
          L.fu1_8.alloc((size_t)m.nfe * m.FF); L.fv1_8.alloc((size_t)m.nfe * m.FF);
          L.fuq_8.alloc((size_t)m.nfe * 3 * D); L.fvq_8.alloc((size_t)m.nfe * 3 * D);
          fold_vectors(L.w8_ff1.p, D, D, 2, L.mod.p + 4 * D, L.mod.p + 3 * D, 6 * D, L.ff1.bias.p, L.fu1_8.p, L.fv1_8.p, m.FF, m.nfe, s, L.sw_ff1.p);
          fold_vectors(L.w8_qkv.p, D, D, 2, L.mod.p + D, L.mod.p, 6 * D, L.qkv.bias.p, L.fuq_8.p, L.fvq_8.p, 3 * D, m.nfe, s, L.sw_qkv.p);
          // level 2 (ff2): weights [D][FF], the hidden activation is written as 16 * GELU(.) in e4m3 (dit_chain.cu: CH_FP8_HGAIN)
          L.w8_ff2.alloc((size_t)D * m.FF); L.sw_ff2.alloc(D); L.gate8.alloc((size_t)m.nfe * D); L.bias8.alloc(D);
          quantize_rows_e4m3(L.ff2.w_ref, D, m.FF, L.w8_ff2.p, m.FF, L.sw_ff2.p, s);
          fold_gate_bias(L.mod.p + 5 * D, 6 * D, L.ff2.bias.p, L.sw_ff2.p, 16.0f, L.gate8.p, L.bias8.p, D, m.nfe, s);
          L.fp8_ready = true;
        }
      }
    }
  }
  // q | k | v of block l from the LN-modulated rows in `a16`: one GEMM, RoPE + V^T in the epilogue (modules.py:459-466)
  auto qkv_fast = [&](DiTLayer& L, const __nv_bfloat16* a16) {
    RowGemm p;
    p.x = a16; p.ldx = D; p.Lin = R; p.Cin = D; p.N = 3 * D; p.taps = 1; p.M = R; p.B = 1;
    p.out = m.qk16.p; p.ldo = 2 * D; p.out_bf16 = fast; p.o_limit = (long)R * 2 * D + 3 * D;
    p.bias = L.qkv.bias.p; p.f16 = f16;
    p.rope_cs = m.rope_cs16.p; p.rope_cols = 2 * D; p.rope_rows = N; p.rowinfo = rowinfo;
    p.vt_out = m.vT16.p; p.vt_col0 = 2 * D; p.vt_ld = m.Npad; p.vt_heads = m.H;
    ProfScope ps(e.prof, "f5.qkv_gemm", s);
    rowgemm_tc(p, L.qkv.tc[f16], s);
  };
  const size_t flag_words = dit_chain_flag_words(R);
  // debug: B200TTS_CHAIN_TRACE=<file> (with B200TTS_GRAPHS=0) dumps the %globaltimer stamps of the last-but-one block's chain launch, [CTA][64] u64
  const char* trace_path = chain ? getenv("B200TTS_CHAIN_TRACE") : nullptr;
  if (trace_path) { m.chain_trace.reserve((size_t)512 * 64); B2_CUDA(cudaMemsetAsync(m.chain_trace.p, 0, (size_t)512 * 64 * 8, s)); }
  for (int step = first; step < first + count; ++step) {
    // ---- input embedding: h[b] = Wx x + (Wc c_b + bias) ; x = conv_pos(h) + h ----
    if (fast) {
      // tensor-core form: x of all U utterances as one batched A operand (16-bit, rows padded to 8), one launch per CFG row
      // (both rows share x); the epilogue adds the step-invariant half and also writes the 16-bit copy conv_pos reads
      const int ldx = (int)round_up(m.n_mels, 8);
      // ragged batch, option ragged_embed: the 1-tap embedding GEMM does not care where a sequence ends, so x is gathered into
      // the DiT row order (both CFG sequences of an utterance read its tokens) and ALL 2 * Ntot rows are one launch instead of two
      // per utterance; the grouped convolutions below keep their per-utterance launches (a sequence's length is their batch stride)
      const bool one_embed = m.ragged && e.ragged_embed;
      lin_prepare_tc(e, m.wx, f16);
      if (one_embed) {
        { ProfScope ps(e.prof, "f5.cast", s); cast_pad_rows_ragged(m.noise.p, m.x16.p, m.rowinfo.p, m.seq_off.p, R, m.n_mels, ldx, s, f16); }
        RowGemm p;
        p.x = m.x16.p; p.x_bstride = (long)R * ldx; p.ldx = ldx; p.Lin = R;
        p.Cin = m.n_mels; p.N = D; p.taps = 1; p.M = R; p.B = 1;
        p.out = m.h.p; p.o_bstride = (long)R * D; p.ldo = D;
        p.res = m.cproj.p;
        p.out2 = m.h16.p; p.f16 = f16;
        ProfScope ps(e.prof, "f5.embed_x", s);
        rowgemm_tc(p, m.wx.tc[f16], s);
      } else {
        ProfScope ps(e.prof, "f5.cast", s); cast_pad_f32_to_bf16(m.noise.p, m.x16.p, m.Ntot, m.n_mels, ldx, s, f16);
      }
      // uniform batch: one launch per CFG row over all utterances and one grouped conv over all 2U sequences; ragged batch:
      // the same per utterance (its two sequences), since a sequence's length is the batch stride of these launches
      const int groups = m.ragged ? m.U : 1;
      for (int gi = 0; gi < groups; ++gi) {
        const int Ng = m.ragged ? m.Nu[gi] : N, Ug = m.ragged ? 1 : m.U;
        const size_t t0 = m.ragged ? (size_t)m.tok_off[gi] : 0, r0 = 2 * t0;
        for (int b = 0; b < 2 && !one_embed; ++b) {
          RowGemm p;
          p.x = m.x16.p + t0 * ldx; p.x_bstride = (long)Ng * ldx; p.ldx = ldx; p.Lin = Ng;
          p.Cin = m.n_mels; p.N = D; p.taps = 1; p.M = Ng; p.B = Ug;
          p.out = m.h.p + (r0 + (size_t)b * Ng) * D; p.o_bstride = (long)2 * Ng * D; p.ldo = D;
          p.res = m.cproj.p + (r0 + (size_t)b * Ng) * D;
          p.out2 = m.h16.p + (r0 + (size_t)b * Ng) * D; p.f16 = f16;
          ProfScope ps(e.prof, "f5.embed_x", s);
          rowgemm_tc(p, m.wx.tc[f16], s);
        }
        gconv(e, m.cp1, fast, m.h16.p + r0 * D, Ng, 2 * Ug, m.c16.p + r0 * D, fast, ACT_MISH, nullptr);
        gconv(e, m.cp2, fast, m.c16.p + r0 * D, Ng, 2 * Ug, m.x.p + r0 * D, 0, ACT_MISH, m.h.p + r0 * D);
      }
    } else {
      for (int sq = 0; sq < S; ++sq) {           // sequence sq = (utterance sq/2, CFG row sq%2): both rows share x
        Epi ep; ep.res = m.cproj.p + (size_t)sq * N * D;
        linear(e, "f5.embed_x", m.wx, 0, m.noise.p + (size_t)(sq / 2) * N * m.n_mels, m.n_mels, N, m.h.p + (size_t)sq * N * D, D, ep);
      }
      gconv(e, m.cp1, 0, m.h.p, N, S, m.c32.p, 0, ACT_MISH, nullptr);
      gconv(e, m.cp2, 0, m.c32.p, N, S, m.x.p, 0, ACT_MISH, m.h.p);
    }
    const float* mf = m.mod_final.p + (size_t)step * 2 * D;       // chunk order: scale, shift (modules.py:323)
    if (chain) {
      // ---- 22 DiT blocks, fused: block 0's LN + q|k|v as two launches, then per block attention + ONE chain kernel that runs
      //      out-proj, LN, ff1, ff2, LN and the NEXT block's q|k|v (dit_chain.cu) ----
      B2_CUDA(cudaMemsetAsync(m.chain_flags.p, 0, flag_words * m.depth * sizeof(unsigned), s));
      {
        const float* mod0 = m.layers[0].mod.p + (size_t)step * 6 * D;
        { ProfScope ps(e.prof, "f5.ln_modulate", s); ln_modulate(m.x.p, mod0 + D, mod0, m.n16b.p, fast, R, D, s, m.rowscale.p); }
        qkv_fast(m.layers[0], m.n16b.p);
      }
      for (int l = 0; l < m.depth; ++l) {
        DiTLayer& L = m.layers[l];
        const float* mod = L.mod.p + (size_t)step * 6 * D;
        { ProfScope ps(e.prof, "f5.attention", s); attention_tc(m.qk16.p, m.vT16.p, m.Npad, m.att16.p, S, N, m.H, s, f16, m.ragged ? m.seq_off.p : nullptr, m.ragged ? m.seq_len.p : nullptr, R); }
        const bool last = l + 1 == m.depth;
        const float* nxt = last ? nullptr : m.layers[l + 1].mod.p + (size_t)step * 6 * D;
        DitChain c;
        c.R = R; c.D = D; c.FF = m.FF; c.f16 = f16; c.has_qkv = last ? 0 : 1;
        c.att16 = m.att16.p; c.x = m.x.p; c.n16 = m.n16.p; c.ff16 = m.ff16.p; c.n16b = m.n16b.p;
        c.w_out = &L.out.tc[f16]; c.w_ff1 = &L.ff1.tc[f16]; c.w_ff2 = &L.ff2.tc[f16];
        c.b_out = L.out.bias.p; c.gate_msa = mod + 2 * D; c.shift_mlp = mod + 3 * D; c.scale_mlp = mod + 4 * D;
        c.b_ff1 = L.ff1.bias.p; c.b_ff2 = L.ff2.bias.p; c.gate_mlp = mod + 5 * D;
        if (last) { c.scale_nxt = mf; c.shift_nxt = mf + D; }
        else {
          c.shift_nxt = nxt; c.scale_nxt = nxt + D;
          c.w_qkv = &m.layers[l + 1].qkv.tc[f16]; c.b_qkv = m.layers[l + 1].qkv.bias.p;
          c.qk16 = m.qk16.p; c.rope_cs = m.rope_cs16.p; c.rope_rows = N; c.rowinfo = rowinfo; c.vt_out = m.vT16.p; c.vt_ld = m.Npad; c.vt_heads = m.H;
        }
        if (e.dit_fp8) {
          c.fp8 = e.dit_fp8 >= 2 ? 2 : 1;
          if (c.fp8 == 2) { c.w8_ff2 = L.w8_ff2.p; c.gate_mlp = L.gate8.p + (size_t)step * D; c.b_ff2 = L.bias8.p; }
          c.w8_ff1 = L.w8_ff1.p; c.sw_ff1 = L.sw_ff1.p;
          c.u_ff1 = L.fu1_8.p + (size_t)step * m.FF; c.v_ff1 = L.fv1_8.p + (size_t)step * m.FF;
          if (!last) {
            DiTLayer& Ln = m.layers[l + 1];
            c.w8_qkv = Ln.w8_qkv.p; c.sw_qkv = Ln.sw_qkv.p;
            c.u_qkv = Ln.fuq_8.p + (size_t)step * 3 * D; c.v_qkv = Ln.fvq_8.p + (size_t)step * 3 * D;
          }
        } else {
          c.u_ff1 = L.fu1[f16].p + (size_t)step * m.FF; c.v_ff1 = L.fv1[f16].p + (size_t)step * m.FF;
          if (!last) { c.u_qkv = m.layers[l + 1].fuq[f16].p + (size_t)step * 3 * D; c.v_qkv = m.layers[l + 1].fvq[f16].p + (size_t)step * 3 * D; }
        }
        c.rowscale = m.rowscale.p;
        c.stats = m.chain_stats.p; c.flags = m.chain_flags.p + (size_t)l * flag_words;
        c.trace = trace_path && l == m.depth - 2 ? m.chain_trace.p : nullptr;      // a typical block (the last one has no q|k|v job)
        ProfScope ps(e.prof, "f5.chain", s);
        dit_chain(c, s);
      }
    } else {
    // ---- 22 DiT blocks ----
    for (int l = 0; l < m.depth; ++l) {
      DiTLayer& L = m.layers[l];
      const float* mod = L.mod.p + (size_t)step * 6 * D;
      const float *shift_msa = mod, *scale_msa = mod + D, *gate_msa = mod + 2 * D;
      const float *shift_mlp = mod + 3 * D, *scale_mlp = mod + 4 * D, *gate_mlp = mod + 5 * D;
      void* nbuf = fast ? (void*)m.n16.p : (void*)m.n32.p;
      { ProfScope ps(e.prof, "f5.ln_modulate", s); ln_modulate(m.x.p, scale_msa, shift_msa, nbuf, fast, R, D, s); }
      if (fast) {
        qkv_fast(L, m.n16.p);
        { ProfScope ps(e.prof, "f5.attention", s); attention_tc(m.qk16.p, m.vT16.p, m.Npad, m.att16.p, S, N, m.H, s, f16, m.ragged ? m.seq_off.p : nullptr, m.ragged ? m.seq_len.p : nullptr, R); }
      } else {
        Epi ep; ep.bias = L.qkv.bias.p;
        linear(e, "f5.qkv_gemm", L.qkv, 0, m.n32.p, D, R, m.qkv32.p, 3 * D, ep);
        {
          ProfScope ps(e.prof, "f5.rope_split", s);
          rope_split_f32(m.qkv32.p, m.cur_cos, m.cur_sin, m.kT32.p, m.v32.p, N, m.H, m.hd, m.Npad, s);
        }
        attention_f32(e, m, 0);
        attention_f32(e, m, 1);
      }
      {
        Epi ep; ep.bias = L.out.bias.p; ep.gate = gate_msa; ep.res = m.x.p;      // x += gate_msa * attn
        linear(e, "f5.out_gemm", L.out, fast, fast ? (const void*)m.att16.p : (const void*)m.att32.p, D, R, m.x.p, D, ep);
      }
      { ProfScope ps(e.prof, "f5.ln_modulate", s); ln_modulate(m.x.p, scale_mlp, shift_mlp, nbuf, fast, R, D, s); }
      {
        Epi ep; ep.bias = L.ff1.bias.p; ep.act = ACT_GELU_TANH; ep.out_bf16 = fast;
        linear(e, "f5.ff1_gemm", L.ff1, fast, nbuf, D, R, fast ? (void*)m.ff16.p : (void*)m.ff32.p, m.FF, ep);
      }
      {
        Epi ep; ep.bias = L.ff2.bias.p; ep.gate = gate_mlp; ep.res = m.x.p;      // x += gate_mlp * ff
        linear(e, "f5.ff2_gemm", L.ff2, fast, fast ? (const void*)m.ff16.p : (const void*)m.ff32.p, m.FF, R, m.x.p, D, ep);
      }
    }
    }
    // ---- final modulation, projection, CFG + Euler ----
    void* nbuf = fast ? (chain ? (void*)m.n16b.p : (void*)m.n16.p) : (void*)m.n32.p;
    if (!chain) { ProfScope ps(e.prof, "f5.ln_modulate", s); ln_modulate(m.x.p, mf, mf + D, nbuf, fast, R, D, s); }
    {
      Epi ep; ep.bias = m.proj.bias.p;
      linear(e, "f5.proj_out", m.proj, fast, nbuf, D, R, m.pred.p, m.n_mels, ep);
    }
    {
      ProfScope ps(e.prof, "f5.euler", s);
      if (!m.ragged) {
        euler_cfg_update(m.noise.p, m.pred.p, (long)N * m.n_mels, m.U, m.cfg_strength, m.delta_t[step], s);
      } else {
        for (int u = 0; u < m.U; ++u)
          euler_cfg_update(m.noise.p + (size_t)m.tok_off[u] * m.n_mels, m.pred.p + (size_t)2 * m.tok_off[u] * m.n_mels, (long)m.Nu[u] * m.n_mels, 1,
                           m.cfg_strength, m.delta_t[step], s);
      }
    }
  }
  if (trace_path) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      std::vector<unsigned long long> h((size_t)512 * 64);
      B2_CUDA(cudaMemcpyAsync(h.data(), m.chain_trace.p, h.size() * 8, cudaMemcpyDeviceToHost, s));
      B2_CUDA(cudaStreamSynchronize(s));
      if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    }
  }
}

// =============================================================================================
// graph C
// =============================================================================================
long f5_decode(Engine& e, const float* d_mel, int N, int ref_len, int16_t* d_pcm, float* d_wave, int fast) {
  F5Model& m = model(e);
  cudaStream_t s = e.stream;
  if (d_mel == nullptr) { d_mel = m.noise.p; }
  B2_CHECK(ref_len >= 0 && ref_len < N, "decode: ref_signal_len must be in [0, max_duration)");
  const int G = N - ref_len;
  B2_CHECK(G <= m.max_frames, "decode: more frames than the ISTFT window-sum table covers");
  const int f16 = fast == PREC_F16 ? 1 : 0;
  const float* x0 = d_mel + (size_t)ref_len * m.n_mels;        // slice [:, ref_len:] -- rows are frames
  const int VC = m.VC, VI = m.VI;
  m.d_a.reserve((size_t)G * VC); m.d_b.reserve((size_t)G * VC); m.d_c.reserve((size_t)G * VC);
  m.d_wide.reserve((size_t)G * VI);
  m.d_head.reserve((size_t)G * m.head.Np); m.d_in.reserve((size_t)G * m.istft.K); m.d_frames.reserve((size_t)G * m.nfft);
  {
    RowGemm p;                                  // embed: Conv1d(n_mels -> VC, k 7, pad 3)
    p.x = x0; p.ldx = m.n_mels; p.Lin = G; p.Cin = m.n_mels; p.N = VC; p.taps = 7; p.center = 3; p.M = G; p.B = 1;
    p.w = m.v_embed_w.p; p.ldw = VC; p.out = m.d_a.p; p.ldo = VC; p.bias = m.v_embed_b.p;
    if (fast) {
      if (!m.v_embed_tc[f16].ready) {
        DevBuf<float> tmp((size_t)7 * VC * m.n_mels);
        conv_weight_permute(m.v_embed_ref, tmp.p, VC, m.n_mels, 7, 1, 0, s);
        tc_weight_from_f32(m.v_embed_tc[f16], tmp.p, 1, 7, VC, m.n_mels, s, f16);
        B2_CUDA(cudaStreamSynchronize(s));
      }
      int l16 = 0;
      p.x = stage_bf16(e, m.s16a, x0, m.n_mels, G, m.n_mels, &l16, f16);
      p.ldx = l16; p.f16 = f16;
      ProfScope ps(e.prof, "f5.vocos_gemm", s);
      rowgemm_tc(p, m.v_embed_tc[f16], s);
    } else {
      ProfScope ps(e.prof, "f5.vocos_gemm", s);
      rowgemm_f32(p, s);
    }
  }
  { ProfScope ps(e.prof, "f5.vocos_elementwise", s); l2_norm_affine(m.d_a.p, m.v_nw.p, m.v_nb.p, m.d_b.p, G, VC, s); }
  float* cur = m.d_b.p; float* other = m.d_a.p;
  for (auto& vb : m.vblocks) {
    {
      ProfScope ps(e.prof, "f5.vocos_elementwise", s);
      dwconv7(cur, vb.dw.p, vb.dwb.p, m.d_c.p, 1, G, VC, s);
      l2_norm_affine(m.d_c.p, vb.nw.p, vb.nb.p, m.d_c.p, G, VC, s);
    }
    Epi e1; e1.bias = vb.pw1.bias.p; e1.act = ACT_GELU_ERF;
    Epi e2; e2.bias = vb.pw2.bias.p; e2.res = cur;
    if (fast) {                        // pw1 writes its GELU output as bf16: pw2's A operand, no staging pass
      int l16 = 0;
      const __nv_bfloat16* a1 = stage_bf16(e, m.s16a, m.d_c.p, VC, G, VC, &l16, f16);
      m.s16b.reserve((size_t)G * VI);
      e1.out_bf16 = fast;
      linear(e, "f5.vocos_gemm", vb.pw1, fast, a1, l16, G, m.s16b.p, VI, e1);
      linear(e, "f5.vocos_gemm", vb.pw2, fast, m.s16b.p, VI, G, other, VC, e2);
    } else {
      linear(e, "f5.vocos_gemm", vb.pw1, 0, m.d_c.p, VC, G, m.d_wide.p, VI, e1);
      linear(e, "f5.vocos_gemm", vb.pw2, 0, m.d_wide.p, VI, G, other, VC, e2);
    }
    std::swap(cur, other);
  }
  { ProfScope ps(e.prof, "f5.vocos_elementwise", s); l2_norm_affine(cur, m.v_fw.p, m.v_fb.p, m.d_c.p, G, VC, s); }
  Epi eh; eh.bias = m.head.bias.p;
  if (fast) {
    int l16 = 0;
    const __nv_bfloat16* ah = stage_bf16(e, m.s16a, m.d_c.p, VC, G, VC, &l16, f16);
    if (!m.head_tc[f16].ready) {                     // N = nfft + 2 is not a multiple of 4: zero weight rows up to Np (bias is padded too)
      DevBuf<float> tmp((size_t)m.head.Np * m.head.K);
      B2_CUDA(cudaMemsetAsync(tmp.p, 0, (size_t)m.head.Np * m.head.K * sizeof(float), s));
      B2_CUDA(cudaMemcpyAsync(tmp.p, m.head.w_ref, (size_t)m.head.N * m.head.K * sizeof(float), cudaMemcpyDeviceToDevice, s));
      tc_weight_from_f32(m.head_tc[f16], tmp.p, 1, 1, m.head.Np, m.head.K, s, f16);
      B2_CUDA(cudaStreamSynchronize(s));
    }
    RowGemm p;
    p.x = ah; p.ldx = l16; p.Lin = G; p.Cin = m.head.K; p.N = m.head.Np; p.taps = 1; p.M = G; p.B = 1;
    p.out = m.d_head.p; p.ldo = m.head.Np; p.bias = eh.bias; p.f16 = f16;
    ProfScope ps(e.prof, "f5.vocos_gemm", s);
    rowgemm_tc(p, m.head_tc[f16], s);
  } else {
    linear(e, "f5.vocos_gemm", m.head, 0, m.d_c.p, VC, G, m.d_head.p, m.head.Np, eh);
  }
  { ProfScope ps(e.prof, "f5.vocos_elementwise", s); istft_input(m.d_head.p, m.d_in.p, G, m.bins, m.head.Np, s); }
  B2_CHECK(m.head.Np == m.istft.K, "head / istft padding mismatch");
  Epi ei;
  if (fast) {
    if (!m.istft_tc[f16].ready) {                    // W^T = basis^T: [nfft][2*bins (+ pad)]
      DevBuf<float> tmp((size_t)m.nfft * m.istft.K);
      transpose_pad(m.istft_ref, tmp.p, 2 * m.bins, m.nfft, m.istft.K, s);
      tc_weight_from_f32(m.istft_tc[f16], tmp.p, 1, 1, m.nfft, m.istft.K, s, f16);
      B2_CUDA(cudaStreamSynchronize(s));
    }
    int l16 = 0;
    RowGemm p;
    p.x = stage_bf16(e, m.s16a, m.d_in.p, m.istft.K, G, m.istft.K, &l16, f16);
    p.ldx = l16; p.Lin = G; p.Cin = m.istft.K; p.N = m.nfft; p.taps = 1; p.M = G; p.B = 1;
    p.out = m.d_frames.p; p.ldo = m.nfft; p.f16 = f16;
    ProfScope ps(e.prof, "f5.istft_gemm", s);
    rowgemm_tc(p, m.istft_tc[f16], s);
  } else {
    linear(e, "f5.istft_gemm", m.istft, 0, m.d_in.p, m.istft.K, G, m.d_frames.p, m.nfft, ei);
  }
  B2_CHECK((long)m.nfft + (long)m.hop * (G - 1) <= m.wsi_len, "window_sum_inv table too short");
  { ProfScope ps(e.prof, "f5.vocos_elementwise", s); istft_overlap_add(m.d_frames.p, m.wsi, G, m.nfft, m.hop, d_pcm, d_wave, s); }
  return (long)m.hop * (G - 1);
}

}  // namespace b200tts

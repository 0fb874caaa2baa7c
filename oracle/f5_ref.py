"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement (PyTorch fp32 eager) of the three F5-TTS graphs the reference exports, quirks included:
  F5_TTS/Export_F5.py:98-141    F5Preprocess   (custom STFT -> log-mel, text embed, RoPE tables, noise)
  F5_TTS/Export_F5.py:144-182   F5Transformer  (one NFE step: DiT on the CFG pair, Euler update)
  F5_TTS/Export_F5.py:185-203   F5Decode       (Vocos backbone/head + custom ISTFT -> int16)
  F5_TTS/Export_F5.py:321-333   Q/K pre-scale; :389-402 Vocos weight folding
  F5_TTS/modeling_modified/F5/dit.py:32-87,205-220, modules.py:167-190,196-261,292-340,421-468,599-613,688-698
  F5_TTS/STFT_Process.py:46-166
  F5_TTS/modeling_modified/vocos/models.py:78-83, modules.py:43-51, heads.py:55-59
  host loop: F5_TTS/F5-TTS-ONNX-Inference.py:247-311
Pinned against the reference modules by oracle/ref_harness.py (tests/golden/f5_*.npz).
State dicts use the reference's names (see text-to-speech-tts-onnx_b200/synth.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
import torchaudio


def _t(sd, name):
    v = sd[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


# ----------------------------------------------------------------------------------------------
# STFT / ISTFT as convolutions (STFT_Process.py)
# ----------------------------------------------------------------------------------------------
def stft_kernels(n_fft=1024):
    """STFT_Process.py:87-98 (win_length == n_fft, periodic hann): fp32 angle 2*pi*f*t/n_fft (quirk q8)."""
    window = torch.hann_window(n_fft).float()
    t = torch.arange(n_fft).float().unsqueeze(0)
    f = torch.arange(n_fft // 2 + 1).float().unsqueeze(1)
    omega = 2 * torch.pi * f * t / n_fft
    cos_k = (torch.cos(omega) * window.unsqueeze(0)).unsqueeze(1)
    sin_k = (-torch.sin(omega) * window.unsqueeze(0)).unsqueeze(1)
    return cos_k, sin_k


def stft_B(x, n_fft=1024, hop=256, kernels=None):
    """STFT_Process.py:144-157 with 'reflect' padding: x (1,1,L) -> real, imag (1, n_fft/2+1, L//hop+1)."""
    cos_k, sin_k = kernels or stft_kernels(n_fft)
    xp = F.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    return F.conv1d(xp, cos_k, stride=hop), F.conv1d(xp, sin_k, stride=hop)


def istft_tables(n_fft=1024, hop=256, max_frames=4096):
    """STFT_Process.py:101-133: inverse basis = hann * pinv(F * n_fft / hop)^T and 1/window-sum for max_frames
    (quirk q9: the normaliser assumes 4096 frames)."""
    window = torch.hann_window(n_fft).float()
    half = n_fft // 2
    fb = torch.fft.fft(torch.eye(n_fft, dtype=torch.float32))
    fb = torch.vstack([torch.real(fb[: half + 1]), torch.imag(fb[: half + 1])]).float()
    inverse_basis = window * torch.linalg.pinv((fb * n_fft) / hop).T.unsqueeze(1)        # (n_fft+2, 1, n_fft)
    n = n_fft + hop * (max_frames - 1)
    window_sum = torch.zeros(n, dtype=torch.float32)
    wn = window / window.abs().max()
    win_sq = wn ** 2
    for i in range(max_frames):
        s = i * hop
        window_sum[s:s + n_fft] += win_sq[: max(0, min(n_fft, n - s))]
    window_sum_inv = n_fft / (window_sum * hop + 1e-7)
    return inverse_basis, window_sum_inv


def istft_A(mag, phase, tables, n_fft=1024, hop=256):
    """STFT_Process.py:160-166."""
    inverse_basis, window_sum_inv = tables
    inp = torch.cat((mag * torch.cos(phase), mag * torch.sin(phase)), dim=1)
    inv = F.conv_transpose1d(inp, inverse_basis, stride=hop)
    s, e = n_fft // 2, inv.size(-1) - n_fft // 2
    return inv[:, :, s:e] * window_sum_inv[s:e]


# ----------------------------------------------------------------------------------------------
# constants of graph A / B
# ----------------------------------------------------------------------------------------------
def mel_fbank(cfg):
    """Export_F5.py:113: HTK, no norm, 0..sr/2 -> (1, n_mels, n_fft/2+1)."""
    fb = torchaudio.functional.melscale_fbanks(cfg.nfft // 2 + 1, 0, cfg.sample_rate // 2, cfg.n_mels, cfg.sample_rate, None, "htk")
    return fb.transpose(0, 1).unsqueeze(0)


def rope_tables(cfg):
    """Export_F5.py:107-112: theta 10000, interleave-repeated, rounded through fp16 (quirk q5) -> (max_frames, head_dim)."""
    hd = cfg.head_dim
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))
    freqs = torch.outer(torch.arange(cfg.max_frames, dtype=torch.float32), inv_freq)
    freqs = freqs.repeat_interleave(2, dim=-1)
    return freqs.cos().half().float(), freqs.sin().half().float()


def text_pos_table(cfg, max_pos=4096):
    """modules.py:196-207 precompute_freqs_cis(text_dim, 4096): [cos | sin] (max_pos, text_dim)."""
    dim = cfg.text_dim
    freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    freqs = torch.outer(torch.arange(max_pos), freqs).float()
    return torch.cat([torch.cos(freqs), torch.sin(freqs)], dim=-1)


def time_tables(sd, cfg):
    """Export_F5.py:153-164: sway-sampled grid, delta_t (nfe-1,), time_expand (nfe, dim) = time_mlp(sin|cos)."""
    t = torch.linspace(0, 1, cfg.nfe, dtype=torch.float32)
    time_step = t + cfg.sway * (torch.cos(torch.pi * 0.5 * t) - 1 + t)
    delta_t = torch.diff(time_step)
    half = 128
    k = math.log(10000) / (half - 1)
    fac = 1000.0 * torch.exp(torch.arange(half, dtype=torch.float32) * -k)
    rows = []
    for i in range(cfg.nfe):
        emb = time_step[i] * fac
        emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
        h = F.linear(emb, _t(sd, "time_embed.time_mlp.0.weight"), _t(sd, "time_embed.time_mlp.0.bias"))
        h = F.silu(h)
        rows.append(F.linear(h, _t(sd, "time_embed.time_mlp.2.weight"), _t(sd, "time_embed.time_mlp.2.bias")))
    return delta_t, torch.stack(rows, 0), time_step


def prescale_qk(sd, cfg):
    """Export_F5.py:321-333 (fp32 graph): Wq, bq, Wk, bk *= head_dim ** -0.25. Returns a new dict."""
    s = math.pow(cfg.head_dim, -0.25)
    out = dict(sd)
    for i in range(cfg.depth):
        for nm in ("to_q", "to_k"):
            for wb in ("weight", "bias"):
                k = f"transformer_blocks.{i}.attn.{nm}.{wb}"
                out[k] = (_t(sd, k) * s)
    return out


def fold_vocos(sd, cfg):
    """Export_F5.py:390-402: norm weights x sqrt(C), gamma folded into pwconv2. Returns a new dict of tensors with
    the plain (unfolded-layout) shapes: norm weight (C,), pwconv weights (out, in)."""
    out = {k: _t(sd, k).clone() for k in sd}
    C = cfg.vocos_dim
    rt = torch.sqrt(torch.tensor(C, dtype=torch.float32))
    out["backbone.norm.weight"] = out["backbone.norm.weight"] * rt
    out["backbone.final_layer_norm.weight"] = out["backbone.final_layer_norm.weight"] * rt
    for i in range(cfg.vocos_layers):
        p = f"backbone.convnext.{i}."
        out[p + "norm.weight"] = out[p + "norm.weight"] * rt
        g = out.pop(p + "gamma")
        out[p + "pwconv2.weight"] = g.unsqueeze(-1) * out[p + "pwconv2.weight"]
        out[p + "pwconv2.bias"] = g * out[p + "pwconv2.bias"]
    return out


# ----------------------------------------------------------------------------------------------
# graph A: F5Preprocess
# ----------------------------------------------------------------------------------------------
def _convnext_v2(x, sd, p):
    """modules.py:233-261 + GRN :217-226 on (1, N, D)."""
    res = x
    D = x.shape[-1]
    h = F.conv1d(x.transpose(1, 2), _t(sd, p + "dwconv.weight"), _t(sd, p + "dwconv.bias"), padding=3, groups=D).transpose(1, 2)
    h = F.layer_norm(h, (D,), _t(sd, p + "norm.weight"), _t(sd, p + "norm.bias"), eps=1e-6)
    h = F.linear(h, _t(sd, p + "pwconv1.weight"), _t(sd, p + "pwconv1.bias"))
    h = F.gelu(h)
    gx = torch.norm(h, p=2, dim=1, keepdim=True)                     # L2 over the SEQUENCE dim
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    h = _t(sd, p + "grn.gamma") * (h * nx) + _t(sd, p + "grn.beta") + h
    h = F.linear(h, _t(sd, p + "pwconv2.weight"), _t(sd, p + "pwconv2.bias"))
    return res + h


def text_embed(sd, text, seq_len, cfg):
    """dit.py:49-73: text (1, N) int with 0 = filler. Returns text, text_drop (1, N, text_dim)."""
    emb = _t(sd, "text_embed.text_embed.weight")
    mask = (text == 0).unsqueeze(-1)
    pos = text_pos_table(cfg)[:seq_len].unsqueeze(0)
    outs = []
    for ids in (text, torch.zeros_like(text)):
        h = F.embedding(ids.long(), emb) + pos
        h = h.masked_fill(mask, 0.0)
        for i in range(cfg.text_conv_layers):
            h = _convnext_v2(h, sd, f"text_embed.text_blocks.{i}.")
            h = h.masked_fill(mask, 0.0)                              # text_drop is masked by the REAL text's mask (q11)
        outs.append(h)
    return outs[0], outs[1]


@torch.inference_mode()
def f5_preprocess(audio, text_ids, max_duration, sd, cfg, noise=None):
    """Export_F5.py:117-141 (fp32 graph). audio int16 (1,1,L); text_ids int32 (1,n); max_duration int64 (1,).
    The reference draws ``noise`` with RandomNormalLike inside ORT; pass it explicitly for parity."""
    audio = torch.as_tensor(audio)
    text_ids = torch.as_tensor(text_ids)
    N = int(np.asarray(max_duration).reshape(-1)[0])
    a = audio.float() * float(1.0 / 32768.0)
    real, imag = stft_B(a, cfg.nfft, cfg.hop)
    mel = torch.matmul(mel_fbank(cfg), torch.sqrt(real * real + imag * imag)).transpose(1, 2).clamp(min=1e-5).log()
    ref_len = mel.shape[1]
    zeros = torch.zeros((1, N, cfg.n_mels), dtype=torch.float32)
    mel = torch.cat((mel, zeros[:, :-ref_len]), dim=1)
    if noise is None:
        noise = torch.randn_like(zeros)
    noise = torch.as_tensor(noise).float()
    cos, sin = rope_tables(cfg)
    rope_cos_q = cos[:N].view(1, 1, N, -1).expand(2, cfg.heads, N, cfg.head_dim)
    rope_sin_q = sin[:N].view(1, 1, N, -1).expand(2, cfg.heads, N, cfg.head_dim)
    pad = torch.zeros((1, N - text_ids.shape[-1]), dtype=text_ids.dtype)
    text, text_drop = text_embed(sd, torch.cat((text_ids + 1, pad), dim=-1), N, cfg)
    cat_mel_text = torch.cat((mel, text), dim=-1)
    cat_mel_text_drop = torch.cat((zeros, text_drop), dim=-1)
    return (noise, rope_cos_q, rope_sin_q, rope_cos_q.transpose(-1, -2), rope_sin_q.transpose(-1, -2),
            cat_mel_text, cat_mel_text_drop, ref_len)


# ----------------------------------------------------------------------------------------------
# graph B: F5Transformer
# ----------------------------------------------------------------------------------------------
def _rope_interleaved(x, cos, sin):
    """modules.py:421-438 for the (.., N, 64) layout: pairs (x0,x1) -> (-x1, x0) (quirk q7)."""
    x1 = x[..., 0::2]
    x2 = x[..., 1::2]
    rot = torch.stack((-x2, x1), dim=-1).reshape(x.shape)
    return x * cos + rot * sin


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps=1e-6)


def dit_forward(sd, x, cond, cond_drop, t, cos, sin, cfg, taps=None):
    """dit.py:205-220 with the Q/K-prescaled state dict. x (1,N,100); cond, cond_drop (1,N,612); t (1024,);
    cos/sin (N,64). Returns (2,N,100)."""
    H, hd, D = cfg.heads, cfg.head_dim, cfg.dim
    N = x.shape[1]

    def embed(c):
        h = F.linear(torch.cat((x, c), dim=-1), _t(sd, "input_embed.proj.weight"), _t(sd, "input_embed.proj.bias"))
        g = h.permute(0, 2, 1)
        for n in (0, 2):
            g = F.mish(F.conv1d(g, _t(sd, f"input_embed.conv_pos_embed.conv1d.{n}.weight"),
                                _t(sd, f"input_embed.conv_pos_embed.conv1d.{n}.bias"), padding=cfg.convpos_kernel // 2,
                                groups=cfg.convpos_groups))
        return g.permute(0, 2, 1) + h

    h = torch.cat((embed(cond), embed(cond_drop)), dim=0)
    if taps is not None:
        taps["embed"] = h
    st = F.silu(t)
    for i in range(cfg.depth):
        p = f"transformer_blocks.{i}."
        emb = F.linear(st, _t(sd, p + "attn_norm.linear.weight"), _t(sd, p + "attn_norm.linear.bias"))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = torch.chunk(emb, 6, dim=-1)
        n = _ln(h) * (1 + scale_msa) + shift_msa
        q = F.linear(n, _t(sd, p + "attn.to_q.weight"), _t(sd, p + "attn.to_q.bias")).view(2, N, H, hd).transpose(1, 2)
        k = F.linear(n, _t(sd, p + "attn.to_k.weight"), _t(sd, p + "attn.to_k.bias")).view(2, N, H, hd).transpose(1, 2)
        v = F.linear(n, _t(sd, p + "attn.to_v.weight"), _t(sd, p + "attn.to_v.bias")).view(2, N, H, hd).transpose(1, 2)
        q = _rope_interleaved(q, cos, sin)
        k = _rope_interleaved(k, cos, sin)
        a = torch.softmax(torch.matmul(q, k.transpose(-1, -2)), dim=-1, dtype=torch.float32)   # no scale (folded), no mask
        a = torch.matmul(a, v).transpose(1, 2).reshape(2, N, D)
        a = F.linear(a, _t(sd, p + "attn.to_out.0.weight"), _t(sd, p + "attn.to_out.0.bias"))
        h = h + gate_msa * a
        n = _ln(h) * (1 + scale_mlp) + shift_mlp
        f = F.gelu(F.linear(n, _t(sd, p + "ff.ff.0.0.weight"), _t(sd, p + "ff.ff.0.0.bias")), approximate="tanh")
        f = F.linear(f, _t(sd, p + "ff.ff.2.weight"), _t(sd, p + "ff.ff.2.bias"))
        h = h + gate_mlp * f
        if taps is not None and i in (0, cfg.depth - 1):
            taps[f"block{i}"] = h
    emb = F.linear(st, _t(sd, "norm_out.linear.weight"), _t(sd, "norm_out.linear.bias"))
    scale, shift = torch.chunk(emb, 2, dim=-1)
    h = _ln(h) * (1 + scale) + shift
    return F.linear(h, _t(sd, "proj_out.weight"), _t(sd, "proj_out.bias"))


@torch.inference_mode()
def f5_transformer_step(sd, noise, cond, cond_drop, time_step, tables, cfg, cos=None, sin=None, taps=None):
    """Export_F5.py:177-182 with FUSE_NFE = 1. Returns (noise', time_step + 1). ``sd`` must be Q/K-prescaled."""
    delta_t, time_expand, _ = tables
    N = noise.shape[1]
    if cos is None:
        c, s = rope_tables(cfg)
        cos, sin = c[:N], s[:N]
    pred = dit_forward(sd, noise, cond, cond_drop, time_expand[time_step], cos, sin, cfg, taps)
    p0, p1 = pred[0:1], pred[1:2]
    noise = noise + (p0 + (p0 - p1) * cfg.cfg_strength) * delta_t[time_step]
    return noise, time_step + 1


# ----------------------------------------------------------------------------------------------
# graph C: F5Decode
# ----------------------------------------------------------------------------------------------
def _rms_style_norm(x, w, b):
    """vocos/models.py:80,83, modules.py:46: w * x / ||x||_2(dim=C) + b on (1,C,L); w already x sqrt(C) (quirk q4)."""
    return w.view(1, -1, 1) * x / torch.norm(x, p=2, dim=1, keepdim=True) + b.view(1, -1, 1)


def vocos_decode(fsd, mel, cfg, gelu_tanh=False):
    """vocos/pretrained.py:100-114 -> models.py:78-83 -> heads.py:55-59 with the FOLDED dict. mel (1,100,G)."""
    x = F.conv1d(mel, fsd["backbone.embed.weight"], fsd["backbone.embed.bias"], padding=3)
    x = _rms_style_norm(x, fsd["backbone.norm.weight"], fsd["backbone.norm.bias"])
    for i in range(cfg.vocos_layers):
        p = f"backbone.convnext.{i}."
        r = x
        x = F.conv1d(x, fsd[p + "dwconv.weight"], fsd[p + "dwconv.bias"], padding=3, groups=x.shape[1])
        x = _rms_style_norm(x, fsd[p + "norm.weight"], fsd[p + "norm.bias"])
        x = torch.matmul(fsd[p + "pwconv1.weight"], x) + fsd[p + "pwconv1.bias"].view(1, -1, 1)
        x = F.gelu(x, approximate="tanh" if gelu_tanh else "none")           # quirk q12: ORT may swap in tanh
        x = torch.matmul(fsd[p + "pwconv2.weight"], x) + fsd[p + "pwconv2.bias"].view(1, -1, 1)
        x = r + x
    x = _rms_style_norm(x, fsd["backbone.final_layer_norm.weight"], fsd["backbone.final_layer_norm.bias"])
    x = torch.matmul(fsd["head.out.weight"], x) + fsd["head.out.bias"].view(1, -1, 1)
    mag, p = x.chunk(2, dim=1)
    return torch.clip(torch.exp(mag), max=1e2), p


@torch.inference_mode()
def f5_decode(denoised, ref_signal_len, fsd, cfg, tables=None, return_float=False):
    """Export_F5.py:197-203: slice off the reference frames, Vocos, ISTFT, clamp, x32767, truncating int16 cast."""
    tables = tables or istft_tables(cfg.nfft, cfg.hop, cfg.max_frames)
    d = torch.as_tensor(denoised).float()[:, int(ref_signal_len):]
    mag, phase = vocos_decode(fsd, d.transpose(1, 2), cfg)
    sig = istft_A(mag, phase, tables, cfg.nfft, cfg.hop)
    y = sig.clamp(min=-1.0, max=1.0) * 32767.0
    pcm = y.to(torch.int16)
    return (pcm, y) if return_float else pcm


# ----------------------------------------------------------------------------------------------
# the host loop (F5-TTS-ONNX-Inference.py:247-311)
# ----------------------------------------------------------------------------------------------
@torch.inference_mode()
def f5_synthesize(audio, text_ids, max_duration, noise, dit_sd, vocos_sd, cfg, steps=None, return_mel=False):
    sd = prescale_qk(dit_sd, cfg)
    tables = time_tables(sd, cfg)
    noise, cq, sq, _, _, cond, cond_drop, ref_len = f5_preprocess(audio, text_ids, max_duration, sd, cfg, noise)
    N = noise.shape[1]
    cos, sin = cq[0, 0], sq[0, 0]
    ts = 0
    for _ in range(cfg.nfe - 1 if steps is None else steps):      # quirk q10: NFE-1 steps over the whole sequence
        noise, ts = f5_transformer_step(sd, noise, cond, cond_drop, ts, tables, cfg, cos, sin)
    pcm = f5_decode(noise, ref_len, fold_vocos(vocos_sd, cfg), cfg)
    return (pcm, noise, ref_len) if return_mel else pcm

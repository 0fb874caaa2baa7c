"""Checkpoint -> engine tensors: the role the ``.onnx`` initializers play in the reference.

Takes the reference's own state dicts (names as ``state_dict()`` gives them after weight-norm removal /
EMA selection) and applies the export-time transforms at the same place the reference does:
  * BigVGAN/Export_BigVGAN.py:53-57     weight norm removed, fp32; the Kaiser-sinc anti-alias filter is a
                                        registered buffer of the model (filter.py:30-62, resample.py:24-26)
  * F5_TTS/Export_F5.py:321-333         Wq, bq, Wk, bk pre-scaled by head_dim**-0.25
  * F5_TTS/Export_F5.py:153-164         sway-sampled time grid, delta_t, time_expand (time_mlp of 32 constants)
  * F5_TTS/Export_F5.py:389-402         Vocos gamma folded into pwconv2, norm weights x sqrt(C)
"""
import math

import numpy as np
import torch


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> np.ndarray:
    """The anti-alias FIR exactly as the reference model builds it at construction time
    (BigVGAN/modeling_modified/filter.py:30-62), so the taps are bit-identical to its ``filter`` buffer."""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = (torch.arange(-half_size, half_size) + 0.5) if even else (torch.arange(kernel_size) - half_size)
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    filt = filt / filt.sum()
    return filt.to(torch.float32).numpy()


def _np(v):
    return np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float32)


def f5_export_constants(dit_state: dict, cfg) -> dict:
    """The constants Export_F5.py bakes into graphs A/B/C, built with the same torch ops on the host so they are
    bit-identical to the reference's: rope tables rounded through fp16 (:107-112), HTK mel filterbank (:113),
    sway-sampled time grid + time_mlp rows (:153-164), text sinus table (modules.py:196-207), the STFT basis with its
    fp32 angle (STFT_Process.py:87-98) and the ISTFT pinv basis / 4096-frame window-sum (STFT_Process.py:101-133)."""
    import torchaudio
    import torch.nn.functional as F
    out = {}
    hd = cfg.head_dim
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))
    freqs = torch.outer(torch.arange(cfg.max_frames, dtype=torch.float32), inv_freq).repeat_interleave(2, dim=-1)
    out["rope_cos"] = freqs.cos().half().float().numpy()
    out["rope_sin"] = freqs.sin().half().float().numpy()
    td = cfg.text_dim
    tf = 1.0 / (10000.0 ** (torch.arange(0, td, 2)[: td // 2].float() / td))
    tf = torch.outer(torch.arange(cfg.max_frames), tf).float()
    out["text_pos"] = torch.cat([torch.cos(tf), torch.sin(tf)], dim=-1).numpy()
    out["fbank"] = torchaudio.functional.melscale_fbanks(cfg.nfft // 2 + 1, 0, cfg.sample_rate // 2, cfg.n_mels,
                                                         cfg.sample_rate, None, "htk").transpose(0, 1).contiguous().numpy()
    # time grid / time_mlp
    t = torch.linspace(0, 1, cfg.nfe, dtype=torch.float32)
    ts = t + cfg.sway * (torch.cos(torch.pi * 0.5 * t) - 1 + t)
    out["delta_t"] = torch.diff(ts).numpy()
    half = 128
    fac = 1000.0 * torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    w0, b0 = (torch.from_numpy(_np(dit_state[f"time_embed.time_mlp.0.{k}"])) for k in ("weight", "bias"))
    w2, b2 = (torch.from_numpy(_np(dit_state[f"time_embed.time_mlp.2.{k}"])) for k in ("weight", "bias"))
    rows = []
    for i in range(cfg.nfe):
        emb = ts[i] * fac
        emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
        rows.append(F.linear(F.silu(F.linear(emb, w0, b0)), w2, b2))
    out["time_expand"] = torch.stack(rows, 0).numpy()
    # STFT / ISTFT bases
    n = cfg.nfft
    window = torch.hann_window(n).float()
    tt = torch.arange(n).float().unsqueeze(0)
    ff = torch.arange(n // 2 + 1).float().unsqueeze(1)
    omega = 2 * torch.pi * ff * tt / n
    out["stft_basis"] = torch.cat([torch.cos(omega) * window.unsqueeze(0), -torch.sin(omega) * window.unsqueeze(0)], 0).numpy()
    fb = torch.fft.fft(torch.eye(n, dtype=torch.float32))
    fb = torch.vstack([torch.real(fb[: n // 2 + 1]), torch.imag(fb[: n // 2 + 1])]).float()
    out["istft_basis"] = (window * torch.linalg.pinv((fb * n) / cfg.hop).T).contiguous().numpy()       # (n+2, n)
    total = n + cfg.hop * (cfg.max_frames - 1)
    wsum = torch.zeros(total, dtype=torch.float32)
    win_sq = (window / window.abs().max()) ** 2
    for i in range(cfg.max_frames):
        s = i * cfg.hop
        wsum[s:s + n] += win_sq[: max(0, min(n, total - s))]
    out["window_sum_inv"] = (n / (wsum * cfg.hop + 1e-7)).numpy()
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


_rope_cache = {}


def f5_rope_rows(cfg):
    """(cos, sin) tables (max_frames, head_dim) fp32, rounded through fp16 as Export_F5.py:107-112 does."""
    key = (cfg.head_dim, cfg.max_frames)
    if key not in _rope_cache:
        hd = cfg.head_dim
        inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))
        freqs = torch.outer(torch.arange(cfg.max_frames, dtype=torch.float32), inv_freq).repeat_interleave(2, dim=-1)
        _rope_cache[key] = (freqs.cos().half().float().numpy(), freqs.sin().half().float().numpy())
    return _rope_cache[key]


def dit_engine_tensors(state: dict, cfg) -> dict:
    """Reference DiT (EMA) state dict -> "dit.*" tensors with Wq, bq, Wk, bk pre-scaled by head_dim**-0.25
    (Export_F5.py:321-333, fp32 graph)."""
    s = np.float32(math.pow(cfg.head_dim, -0.25))
    out = {}
    for k, v in state.items():
        v = _np(v)
        if ".attn.to_q." in k or ".attn.to_k." in k:
            v = (torch.from_numpy(v) * float(s)).numpy()
        if v.ndim > 4:
            raise ValueError(k)
        if k.endswith("grn.gamma") or k.endswith("grn.beta"):
            v = v.reshape(-1)
        out[k] = v
    return out


def vocos_engine_tensors(state: dict, cfg) -> dict:
    """vocos-mel-24khz state dict -> "vocos.*" tensors with the Export_F5.py:390-402 folding applied
    (norm weights x sqrt(C), gamma folded into pwconv2 weight and bias; gamma itself dropped)."""
    t = {k: torch.from_numpy(_np(v)) for k, v in state.items()}
    rt = torch.sqrt(torch.tensor(cfg.vocos_dim, dtype=torch.float32))
    out = {}
    for k, v in t.items():
        if k.endswith(".gamma"):
            continue
        if k in ("backbone.norm.weight", "backbone.final_layer_norm.weight") or (k.endswith(".norm.weight") and ".convnext." in k):
            v = v * rt
        if k.endswith(".pwconv2.weight"):
            v = t[k.replace("pwconv2.weight", "gamma")].unsqueeze(-1) * v
        if k.endswith(".pwconv2.bias"):
            v = t[k.replace("pwconv2.bias", "gamma")] * v
        out[k] = v.contiguous().numpy()
    return out


def ivgan_engine_tensors(state: dict, cfg) -> dict:
    """IndexTTS_F vocoder state (IndexTTS bigvgan generator names + final_norm.*) -> tensors for
    ``Engine.load_state('ivgan', ...)``. Tensors that belong to the conditioning branch the exported graph does not run
    (speaker_encoder.*, cond_layer.*, conds.*; their outputs arrive as graph inputs, Export_IndexTTS.py:497-520) are
    dropped; the strides are added because a ConvTranspose1d weight does not carry them."""
    keep = {k: v for k, v in state.items() if not k.startswith(("speaker_encoder.", "cond_layer.", "conds.", "logit_scale"))}
    out = bigvgan_engine_tensors(keep)
    out["upsample_rates"] = np.asarray(cfg.upsample_rates, dtype=np.float32)
    return out


def igpt_engine_tensors(state: dict, cfg) -> dict:
    """IndexTTS GPT state (gpt.* names as graphs B-E read them: text_embedding, text_pos_embedding.emb, mel_embedding,
    mel_pos_embedding.emb, h.<i>.* of the Hugging Face GPT2Model inside inference_model, ln_f, final_norm, mel_head) -> tensors for
    ``Engine.load_state('igpt', ...)``. Weights stay in the checkpoint's layout and UNscaled; the engine applies the export's q/k
    pre-scale (Export_IndexTTS.py:250-255). ``meta`` carries the loop constants of Inference_IndexTTS_ONNX.py:36-39."""
    out = {}
    for k, v in state.items():
        if k.endswith((".attn.bias", ".attn.masked_bias")):       # GPT2Attention's causal-mask buffers: unused by graph E
            continue
        out[k] = np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float32)
    out["meta"] = np.asarray([cfg.start_text, cfg.stop_text, cfg.start_mel, cfg.stop_mel, cfg.max_generate, cfg.penalty_range,
                              cfg.repeat_penalty, cfg.ln_eps], dtype=np.float32)
    return out


def bigvgan_engine_tensors(state: dict) -> dict:
    """Reference BigVGAN state dict -> tensors for ``Engine.load_state('bigvgan', ...)``."""
    out = {}
    filt = None
    for k, v in state.items():
        v = np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float32)
        if k.endswith(".filter"):
            filt = v.reshape(-1) if filt is None else filt      # every Activation1d carries the same taps
            continue
        out[k] = v
    if filt is None:
        filt = kaiser_sinc_filter1d(0.5 / 2, 0.6 / 2, 12)
    assert filt.shape == (12,)
    out["aa_filter"] = filt.astype(np.float32)
    return out

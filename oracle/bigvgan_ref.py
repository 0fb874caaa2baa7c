"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's BigVGAN graph (mel -> int16 PCM), PyTorch fp32 eager.
Follows, quirks included:
  BigVGAN/Export_BigVGAN.py:37-49            BIGVGAN wrapper (x32767, clamp, truncating int16 cast)
  BigVGAN/modeling_modified/bigvgan.py:384-410   BigVGAN.forward
  BigVGAN/modeling_modified/bigvgan.py:132-140   AMPBlock1.forward
  BigVGAN/modeling_modified/bigvgan.py:359-382   per-stage pad tables (index -1 = the 15-sample pads)
  BigVGAN/modeling_modified/act.py:25-29         Activation1d.forward
  BigVGAN/modeling_modified/resample.py:11-34    UpSample1d (zero pad by concat, x2 gain, [15:-15])
  BigVGAN/modeling_modified/filter.py:30-62,94-98  kaiser_sinc_filter1d, LowPassFilter1d.forward
  SnakeBeta: upstream NVIDIA/BigVGAN activations.py (not vendored); the same formula is in-repo at
  Qwen_TTS/modeling_modified/modeling_qwen3_tts_tokenizer_v2.py:665-684.
Pinned against the reference modules by oracle/ref_harness.py (tests/golden/bigvgan_*.npz).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """filter.py:30-62 -> (kernel_size,) fp32."""
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
    if even:
        time = torch.arange(-half_size, half_size) + 0.5
    else:
        time = torch.arange(kernel_size) - half_size
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    filt = filt / filt.sum()
    return filt.to(torch.float32)


def aa_filter() -> torch.Tensor:
    """The one 12-tap filter used for both x2 up and x2 down (resample.py:24-26,45-50: cutoff 0.25,
    half-width 0.3)."""
    return kaiser_sinc_filter1d(0.5 / 2, 0.6 / 2, 12)


def snakebeta(x, alpha_log, beta_log):
    """x + 1/(exp(beta)+1e-9) * sin^2(x*exp(alpha)), per-channel, logscale parameters."""
    a = torch.exp(alpha_log).view(1, -1, 1)
    b = torch.exp(beta_log).view(1, -1, 1)
    return x + (1.0 / (b + 1e-9)) * torch.pow(torch.sin(x * a), 2)


def activation1d(x, alpha_log, beta_log, filt, up_pad=5, down_pad_l=5, down_pad_r=6):
    """act.py:25-29 on (B,C,L). Stage tables: pads 5 / (5,6) -> length L; post table: 15 / (15,15) -> L+30."""
    C = x.shape[1]
    w = filt.view(1, 1, -1).expand(C, 1, -1)
    x = F.pad(x, (up_pad, up_pad))                               # resample.py:31 (zero concat)
    x = 2 * F.conv_transpose1d(x, w, stride=2, groups=C)         # resample.py:32
    x = x[..., 15:-15]                                           # resample.py:33 (pad_left = pad_right = 15)
    x = snakebeta(x, alpha_log, beta_log)
    x = F.pad(x, (down_pad_l, down_pad_r))                       # filter.py:96
    return F.conv1d(x, w, stride=2, groups=C)                    # filter.py:97


def _t(sd, name):
    v = sd[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


def amp_block1(x, sd, r, k, dilations, filt):
    """bigvgan.py:132-140."""
    for m, d in enumerate(dilations):
        p = f"resblocks.{r}."
        xt = activation1d(x, _t(sd, p + f"activations.{2*m}.act.alpha"), _t(sd, p + f"activations.{2*m}.act.beta"), filt)
        xt = F.conv1d(xt, _t(sd, p + f"convs1.{m}.weight"), _t(sd, p + f"convs1.{m}.bias"),
                      dilation=d, padding=(k * d - d) // 2)
        xt = activation1d(xt, _t(sd, p + f"activations.{2*m+1}.act.alpha"), _t(sd, p + f"activations.{2*m+1}.act.beta"), filt)
        xt = F.conv1d(xt, _t(sd, p + f"convs2.{m}.weight"), _t(sd, p + f"convs2.{m}.bias"),
                      dilation=1, padding=(k - 1) // 2)
        x = xt + x
    return x


@torch.inference_mode()
def bigvgan_forward(mel, sd, cfg, taps=None):
    """mel (B, n_mels, T) fp32 -> float waveform (B,1,256T+30) in [-1,1]. bigvgan.py:384-410.
    ``taps`` (optional dict) receives intermediates for debugging/goldens."""
    mel = mel if isinstance(mel, torch.Tensor) else torch.from_numpy(mel)
    filt = aa_filter()
    x = F.conv1d(mel.float(), _t(sd, "conv_pre.weight"), _t(sd, "conv_pre.bias"), padding=3)
    if taps is not None:
        taps["conv_pre"] = x
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.conv_transpose1d(x, _t(sd, f"ups.{i}.0.weight"), _t(sd, f"ups.{i}.0.bias"),
                               stride=u, padding=(k - u) // 2)
        if taps is not None:
            taps[f"up{i}"] = x
        xs = None
        for j, (rk, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            y = amp_block1(x, sd, i * nk + j, rk, dil, filt)
            xs = y if xs is None else xs + y
        x = xs * float(1.0 / nk)
        if taps is not None:
            taps[f"stage{i}"] = x
    pp = cfg.post_pad
    x = activation1d(x, _t(sd, "activation_post.act.alpha"), _t(sd, "activation_post.act.beta"), filt,
                     up_pad=pp, down_pad_l=pp, down_pad_r=pp)        # bigvgan.py:402 with the index -1 tables
    x = F.conv1d(x, _t(sd, "conv_post.weight"), None, padding=3)     # no bias in the v2 config
    return torch.tanh(x)                                             # forced on: Export_BigVGAN.py:21,41


@torch.inference_mode()
def bigvgan_pcm(mel, sd, cfg, return_float=False):
    """Export_BigVGAN.py:44-49: x32767, clamp to [-32768, 32767], cast (truncates toward zero)."""
    w = bigvgan_forward(mel, sd, cfg)
    y = (w * 32767.0).clamp(min=-32768.0, max=32767.0)
    pcm = y.to(torch.int16)
    return (pcm, y) if return_float else pcm

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

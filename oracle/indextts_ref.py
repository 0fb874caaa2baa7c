"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the vocoder half of IndexTTS (graph IndexTTS_F: GPT latent -> int16 PCM), PyTorch fp32 eager.
Follows, quirks included:
  IndexTTS/Export_IndexTTS.py:292-314               IndexTTS_F.forward (final_norm on hidden[:-2], conv_pre + cond_layer vector,
                                                   per stage ups -> + cond_i -> mean of 3 AMPBlock1, activation_post with the
                                                   index -1 (15-sample) tables, conv_post WITH bias, tanh, clamp(+-1)*32767, cast)
  IndexTTS/modeling_modified/models.py:20-42,130-250  AMPBlock1.forward(x, idx), BigVGAN.__init__ (layer shapes)
  IndexTTS/modeling_modified/act.py:24-29, resample.py:24-45, filter.py:85-107   the idx-table anti-aliased activation: the same
                                                   zero-pad / x2 up / SnakeBeta / down arithmetic as BigVGAN's (oracle/bigvgan_ref.py)
Pinned against the reference modules by oracle/ref_harness.py::build_indextts_f (tests/golden/indextts_ref.npz).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .bigvgan_ref import _t, aa_filter, activation1d, amp_block1


@torch.inference_mode()
def indextts_f_forward(hidden, conds, cond_layer, sd, cfg):
    """hidden (S, gpt_dim), conds[i] (1, C_i, 1), cond_layer (1, C0, 1) -> float waveform (1, 1, hop*(S-2)+30) in [-1, 1]."""
    tt = lambda v: v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
    hidden, cond_layer = tt(hidden).float(), tt(cond_layer).float()
    filt = aa_filter()
    latent = F.layer_norm(hidden[:-2].unsqueeze(0), (cfg.gpt_dim,), _t(sd, "final_norm.weight"), _t(sd, "final_norm.bias"),
                          cfg.ln_eps)                                                           # Export_IndexTTS.py:301
    x = F.conv1d(latent.transpose(1, 2), _t(sd, "conv_pre.weight"), _t(sd, "conv_pre.bias"), padding=3) + cond_layer   # :302
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.conv_transpose1d(x, _t(sd, f"ups.{i}.0.weight"), _t(sd, f"ups.{i}.0.bias"), stride=u, padding=(k - u) // 2)
        x = x + tt(conds[i]).float()                                                           # :306-307
        xs = None
        for j, (rk, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            y = amp_block1(x, sd, i * nk + j, rk, dil, filt)
            xs = y if xs is None else xs + y
        x = xs * float(1.0 / nk)                                                                # :311 (inv_num_kernels)
    pp = cfg.post_pad
    x = activation1d(x, _t(sd, "activation_post.act.alpha"), _t(sd, "activation_post.act.beta"), filt,
                     up_pad=pp, down_pad_l=pp, down_pad_r=pp)                                   # activation_post(latent, -1)
    x = F.conv1d(x, _t(sd, "conv_post.weight"), _t(sd, "conv_post.bias"), padding=3)
    return torch.tanh(x)


@torch.inference_mode()
def indextts_f_pcm(hidden, conds, cond_layer, sd, cfg, return_float=False):
    """Export_IndexTTS.py:313-314: clamp to [-1, 1], x32767, cast (truncates toward zero)."""
    w = indextts_f_forward(hidden, conds, cond_layer, sd, cfg)
    y = w.clamp(min=-1.0, max=1.0) * 32767.0
    pcm = y.to(torch.int16)
    return (pcm, y) if return_float else pcm

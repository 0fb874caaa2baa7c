"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Build-container only: needs /root/reference.

Loads the reference's own ``modeling_modified`` nn.Module sources *where they lie* under
/root/reference (nothing is copied) through small stubs for the un-installed third-party imports,
so that oracle/*_ref.py can be pinned against the real thing and tests/golden/ can be generated
(oracle/make_golden.py). Never used at run time on the GPU box (the reference is not there).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("B200TTS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "BigVGAN", "modeling_modified"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


# ----------------------------------------------------------------------------------------------
# BigVGAN
# ----------------------------------------------------------------------------------------------
class _SnakeBeta(torch.nn.Module):
    """Upstream NVIDIA/BigVGAN activations.SnakeBeta (not vendored by the reference); formula as in
    Qwen_TTS/modeling_modified/modeling_qwen3_tts_tokenizer_v2.py:665-684."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = torch.nn.Parameter(init * alpha if not alpha_logscale else init.clone())
        self.beta = torch.nn.Parameter(init * alpha if not alpha_logscale else init.clone())
        self.no_div_by_zero = 1e-9

    def forward(self, x):
        alpha = self.alpha.unsqueeze(0).unsqueeze(-1)
        beta = self.beta.unsqueeze(0).unsqueeze(-1)
        if self.alpha_logscale:
            alpha = torch.exp(alpha)
            beta = torch.exp(beta)
        return x + (1.0 / (beta + self.no_div_by_zero)) * torch.pow(torch.sin(x * alpha), 2)


class _AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def load_bigvgan_module():
    d = os.path.join(REF, "BigVGAN", "modeling_modified")
    _stub("activations", SnakeBeta=_SnakeBeta, Snake=_SnakeBeta)
    _stub("utils", init_weights=lambda m, mean=0.0, std=0.01: None,
          get_padding=lambda k, dil=1: int((k * dil - dil) / 2))
    _stub("env", AttrDict=_AttrDict)
    _stub("alias_free_activation")
    _stub("alias_free_activation.torch")
    _load("alias_free_activation.torch.filter", os.path.join(d, "filter.py"))
    _load("alias_free_activation.torch.resample", os.path.join(d, "resample.py"))
    _load("alias_free_activation.torch.act", os.path.join(d, "act.py"))
    return _load("ref_bigvgan", os.path.join(d, "bigvgan.py"))


def build_bigvgan(sd: dict, cfg):
    """Reference BigVGAN generator filled from a synthetic state dict, wrapped exactly as
    BigVGAN/Export_BigVGAN.py:37-57 does (weight norm removed, eval, float, tanh forced)."""
    mod = load_bigvgan_module()
    h = _AttrDict(
        num_mels=cfg.num_mels, upsample_rates=list(cfg.upsample_rates),
        upsample_kernel_sizes=list(cfg.upsample_kernel_sizes),
        upsample_initial_channel=cfg.upsample_initial_channel, resblock="1",
        resblock_kernel_sizes=list(cfg.resblock_kernel_sizes),
        resblock_dilation_sizes=[list(d) for d in cfg.resblock_dilation_sizes],
        activation="snakebeta", snake_logscale=True, use_bias_at_final=False, use_tanh_at_final=False)
    model = mod.BigVGAN(h, use_cuda_kernel=False)
    model.remove_weight_norm()
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    bad = [m for m in missing if not m.endswith("filter")]
    assert not bad and not unexpected, (bad, unexpected)
    model = model.eval().float()

    class BIGVGAN(torch.nn.Module):          # restated from Export_BigVGAN.py:37-49 (script is not importable)
        def __init__(self, bigvgan, use_tanh):
            super().__init__()
            self.bigvgan = bigvgan
            self.bigvgan.use_tanh_at_final = use_tanh
            self.use_tanh = use_tanh

        def forward(self, mel_features):
            w = self.bigvgan(mel_features) * 32767.0
            if self.use_tanh:
                w = w.clamp(min=-32768.0, max=32767.0)
            return w.to(torch.int16)

    return BIGVGAN(model, True)

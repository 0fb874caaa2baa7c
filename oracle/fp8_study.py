"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Build-container only.

Parity study for an fp8 (e4m3) operand mode of the DiT GEMMs (north_star "bf16/fp8"): what would the mel /
PCM error be if the A operand and the weights of a GEMM were rounded to e4m3 (per-row activation scale, per-output-channel
weight scale, fp32 accumulation -- the best case a tcgen05 kind::f8f6f4 kernel could implement), everything else in fp32?
Runs the small golden utterance (N = 130) and, with --full, the BASELINE-size one (N = 1126) through the CPU oracle with fake
quantisation patched into torch.nn.functional.linear for the selected projections, and compares with the unquantised run.

    python oracle/fp8_study.py [--full]

The 16-bit rows of the table use the same machinery with fp16 / bf16 rounding, which calibrates it against the measured
engine numbers (fp16 61.9 dB, bf16 44.0 dB at N = 1126)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402,F401
from b200tts import config, synth  # noqa: E402
from oracle import f5_ref  # noqa: E402

E4M3_MAX = 448.0


def q_e4m3_rows(x):
    """per-row (last dim) scale to the e4m3 range, round, scale back"""
    s = x.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / E4M3_MAX
    return (x / s).to(torch.float8_e4m3fn).to(torch.float32) * s


def q16(x, dt):
    return x.to(dt).to(torch.float32)


class Patch:
    """fake-quantise the operands of F.linear calls whose weight has one of the given shapes"""

    def __init__(self, shapes, mode):
        self.shapes, self.mode, self.orig = set(shapes), mode, F.linear
        self.cache = {}

    def __enter__(self):
        def linear(x, w, b=None):
            if tuple(w.shape) in self.shapes and x.dim() == 3:
                key = (w.data_ptr(), self.mode)
                if key not in self.cache:
                    self.cache[key] = q_e4m3_rows(w) if self.mode == "e4m3" else q16(w, torch.float16 if self.mode == "f16" else torch.bfloat16)
                xq = q_e4m3_rows(x) if self.mode == "e4m3" else q16(x, torch.float16 if self.mode == "f16" else torch.bfloat16)
                return self.orig(xq, self.cache[key], b)
            return self.orig(x, w, b)
        F.linear = linear
        return self

    def __exit__(self, *a):
        F.linear = self.orig


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return float(10.0 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30)))


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def run(cfg, dsd, vsd, inputs, shapes, mode):
    audio, text_ids, maxd, noise = inputs
    with torch.inference_mode():
        if mode is None:
            pcm, mel, ref_len = f5_ref.f5_synthesize(torch.from_numpy(audio), torch.from_numpy(text_ids), torch.from_numpy(maxd), torch.from_numpy(noise),
                                                     dsd, vsd, cfg, return_mel=True)
        else:
            with Patch(shapes, mode):
                pcm, mel, ref_len = f5_ref.f5_synthesize(torch.from_numpy(audio), torch.from_numpy(text_ids), torch.from_numpy(maxd), torch.from_numpy(noise),
                                                         dsd, vsd, cfg, return_mel=True)
    return pcm.numpy().astype(np.float64), mel.numpy(), int(ref_len)


def main():
    full = "--full" in sys.argv
    cfg = config.F5
    dsd, vsd = synth.f5_dit_state(4321), synth.vocos_state(2468)
    D, FF = cfg.dim, cfg.dim * cfg.ff_mult
    qkv, ff1, ff2, out = (D, D), (FF, D), (D, FF), (D, D)      # weight shapes (to_q/k/v and to_out share one: both are selected together)
    cases = [("fp16 operands, all four GEMM families", {qkv, ff1, ff2}, "f16"),
             ("bf16 operands, all four GEMM families", {qkv, ff1, ff2}, "bf16"),
             ("e4m3: ff1 only", {ff1}, "e4m3"),
             ("e4m3: ff1 + q|k|v + out", {ff1, qkv}, "e4m3"),
             ("e4m3: all four GEMM families", {qkv, ff1, ff2}, "e4m3")]
    sizes = [(16384, 20)] + ([(144000, 150)] if full else [])
    for audio_len, n_text in sizes:
        inputs = synth.f5_inputs(1, audio_len=audio_len, n_text=n_text)
        pcm0, mel0, ref_len = run(cfg, dsd, vsd, inputs, None, None)
        N = mel0.shape[1]
        print(f"## N = {N} ({audio_len} samples, {n_text} text ids), 31 Euler steps; reference = the same oracle in fp32", flush=True)
        print("| operand rounding | mel cosine | generated-mel cosine | mel max-abs | PCM SNR (dB) |")
        print("|---|---|---|---|---|")
        for name, shapes, mode in cases:
            pcm, mel, _ = run(cfg, dsd, vsd, inputs, shapes, mode)
            print(f"| {name} | {cosine(mel, mel0):.7f} | {cosine(mel[:, ref_len:], mel0[:, ref_len:]):.7f} | {np.abs(mel - mel0).max():.4f} | {snr_db(pcm0, pcm):.1f} |",
                  flush=True)


if __name__ == "__main__":
    main()

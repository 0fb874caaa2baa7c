"""Upstream checkpoint files -> one engine blob (text-to-speech-tts-onnx_b200/checkpoint.py).

    python tools/convert_checkpoint.py --out engine.b200tts [--f5 model_1250000.safetensors] [--vocos pytorch_model.bin]
                                       [--bigvgan bigvgan_generator.pt] [--indextts-gpt gpt.pth] [--indextts-bigvgan bigvgan_generator.pth]
Applies the reference's export-time transforms (EMA selection, weight-norm removal, Q/K pre-scale, Vocos folding, export constants)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200tts  # noqa: F401,E402
from b200tts import checkpoint, config, weights  # noqa: E402


def _load(path):
    if path.endswith(".safetensors"):
        from safetensors.numpy import load_file
        return load_file(path)
    import torch
    return torch.load(path, map_location="cpu", weights_only=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    for k in ("f5", "vocos", "bigvgan", "indextts-gpt", "indextts-bigvgan"):
        ap.add_argument("--" + k)
    a = ap.parse_args()
    parts = {}
    if a.f5:
        dit = checkpoint.f5_dit_from_checkpoint(_load(a.f5))
        parts["dit"] = weights.dit_engine_tensors(dit, config.F5)
        parts["f5"] = weights.f5_export_constants(dit, config.F5)
    if a.vocos:
        parts["vocos"] = weights.vocos_engine_tensors(checkpoint.vocos_from_checkpoint(_load(a.vocos)), config.F5)
    if a.bigvgan:
        parts["bigvgan"] = weights.bigvgan_engine_tensors(checkpoint.bigvgan_from_checkpoint(_load(a.bigvgan)))
    if a.indextts_bigvgan and not a.indextts_gpt:
        # gpt.final_norm is part of the vocoder graph (Export_IndexTTS.py:301): without it the blob's ivgan part cannot be built
        ap.error("--indextts-bigvgan needs --indextts-gpt (the vocoder graph starts with gpt.final_norm)")
    gpt_sd = checkpoint.indextts_gpt_from_checkpoint(_load(a.indextts_gpt)) if a.indextts_gpt else None
    if gpt_sd is not None:
        parts["igpt"] = weights.igpt_engine_tensors(gpt_sd, config.INDEXTTS_GPT)
    if a.indextts_bigvgan:
        sd = checkpoint.bigvgan_from_checkpoint(_load(a.indextts_bigvgan))
        sd["final_norm.weight"], sd["final_norm.bias"] = gpt_sd["final_norm.weight"], gpt_sd["final_norm.bias"]
        parts["ivgan"] = weights.ivgan_engine_tensors(sd, config.INDEXTTS_VOCODER)
    if not parts:
        ap.error("nothing to convert")
    n = checkpoint.save_blob(a.out, parts)
    print(f"wrote {a.out}: {n / 1e6:.1f} MB, parts {sorted(parts)}")


if __name__ == "__main__":
    main()

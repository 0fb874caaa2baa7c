"""Upstream checkpoints -> engine tensors -> one on-disk blob (the role the `.onnx` initializers play for the
reference). Host-side only; every transform cites the reference line that applies it at export time.

    F5 DiT     F5_TTS/Export_F5.py:204-221   load_checkpoint(..., use_ema=True) of the f5_tts package (un-vendored; published
                                             behaviour: keep the `ema_model.*` entries, drop `initted` / `step` and the mel-spectrogram
                                             buffers) -> strip_ema(); the wrappers only read `.transformer` (:101,147-148)
    BigVGAN    BigVGAN/Export_BigVGAN.py:53-54  BigVGAN.from_pretrained(...).remove_weight_norm() -> remove_weight_norm()
    Vocos      F5_TTS/Export_F5.py:389-402   gamma / sqrt(C) folding is done by weights.vocos_engine_tensors; here only the
                                             feature extractor (unused by graph C) is dropped
    IndexTTS   IndexTTS/Export_IndexTTS.py:319-327  IndexTTS(model_dir).gpt / .bigvgan; key names of the un-vendored index-tts
                                             checkpoints are an ASSUMPTION documented at indextts_gpt_from_checkpoint()

Blob layout (little endian): b"B200TTS1" | uint64 header_bytes | header JSON {"tensors": [{"name", "shape", "offset", "nbytes"}]} padded
to 64 bytes | fp32 data, every tensor 64-byte aligned. `load_blob` memory-maps it; `Engine.load_state` uploads straight from the map.
"""
import json
import os
import struct

import numpy as np

MAGIC = b"B200TTS1"


def _np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().float().numpy()
    return np.asarray(v)


# ----------------------------------------------------------------------------------------------
# transforms
# ----------------------------------------------------------------------------------------------
def strip_ema(state: dict) -> dict:
    """EMA checkpoint of the f5_tts trainer -> model state dict (Export_F5.py:220: use_ema=True)."""
    if "ema_model_state_dict" in state:
        state = state["ema_model_state_dict"]
    drop = ("initted", "step", "mel_spec.mel_stft.mel_scale.fb", "mel_spec.mel_stft.spectrogram.window")
    out = {}
    for k, v in state.items():
        k2 = k[len("ema_model."):] if k.startswith("ema_model.") else k
        if k2 in drop:
            continue
        out[k2] = v
    return out


def remove_weight_norm(state: dict) -> dict:
    """weight = g * v / ||v|| with the norm over every dim but 0 (torch.nn.utils.weight_norm, dim=0), for both spellings:
    `<m>.weight_g` / `<m>.weight_v` and `<m>.parametrizations.weight.original0` / `.original1` (Export_BigVGAN.py:54)."""
    out = {}
    pairs = {}
    for k, v in state.items():
        for g_sfx, v_sfx in ((".weight_g", ".weight_v"), (".parametrizations.weight.original0", ".parametrizations.weight.original1")):
            if k.endswith(g_sfx):
                pairs.setdefault(k[: -len(g_sfx)], {})["g"] = _np(v)
                break
            if k.endswith(v_sfx):
                pairs.setdefault(k[: -len(v_sfx)], {})["v"] = _np(v)
                break
        else:
            out[k] = v
    for base, gv in pairs.items():
        if "g" not in gv or "v" not in gv:
            raise ValueError(f"weight-norm pair of '{base}' is incomplete")
        v = gv["v"].astype(np.float32)
        g = gv["g"].astype(np.float32)
        norm = np.sqrt((v.reshape(v.shape[0], -1).astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
        out[base + ".weight"] = (v * (g.reshape(-1) / norm).reshape((-1,) + (1,) * (v.ndim - 1))).astype(np.float32)
    return out


def f5_dit_from_checkpoint(state: dict) -> dict:
    """model_1250000.safetensors (EMA) -> the DiT state dict weights.dit_engine_tensors / f5_export_constants expect."""
    sd = strip_ema(state)
    out = {k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}
    if not out:                      # already a bare DiT state dict
        out = dict(sd)
    return out


def vocos_from_checkpoint(state: dict) -> dict:
    """charactr/vocos-mel-24khz pytorch_model.bin -> backbone.* / head.* (the mel feature extractor is replaced by graph A)."""
    return {k: v for k, v in state.items() if k.startswith(("backbone.", "head."))}


def bigvgan_from_checkpoint(state: dict) -> dict:
    """nvidia/bigvgan_v2_24khz_100band_256x bigvgan_generator.pt ({'generator': ...}) -> weight-norm-free generator state."""
    sd = state.get("generator", state)
    return remove_weight_norm(sd)


def indextts_gpt_from_checkpoint(state: dict) -> dict:
    """index-tts gpt.pth -> the names graphs B-E read (weights.igpt_engine_tensors). ASSUMED upstream layout (index-tts is not
    vendored; UnifiedVoice of the public repository): `gpt.h.<i>.*` and `gpt.ln_f.*` (the Hugging Face GPT2Model), `text_embedding`,
    `mel_embedding`, `text_pos_embedding.emb`, `mel_pos_embedding.emb`, `final_norm`, `mel_head`, optionally under a top-level
    'model' key. Conditioning-encoder / perceiver tensors (graph A) are dropped."""
    sd = state.get("model", state)
    out = {}
    for k, v in sd.items():
        if k.startswith("gpt.h.") or k.startswith("gpt.ln_f."):
            k2 = k[len("gpt."):]
            if k2.endswith((".attn.bias", ".attn.masked_bias")):
                continue
            out[k2] = v
        elif k.startswith(("text_embedding.", "mel_embedding.", "text_pos_embedding.emb.", "mel_pos_embedding.emb.", "final_norm.", "mel_head.")):
            out[k] = v
    return out


# ----------------------------------------------------------------------------------------------
# blob
# ----------------------------------------------------------------------------------------------
def save_blob(path: str, parts: dict) -> int:
    """parts: {prefix: {name: array}} (what Engine.load_state(prefix, tensors) takes). -> bytes written."""
    entries, arrays = [], []
    off = 0
    for prefix, tensors in parts.items():
        for name, v in tensors.items():
            a = np.ascontiguousarray(_np(v), dtype="<f4")
            entries.append({"name": f"{prefix}.{name}", "shape": list(a.shape), "offset": off, "nbytes": int(a.nbytes)})
            arrays.append(a)
            off += (a.nbytes + 63) // 64 * 64
    header = json.dumps({"tensors": entries}).encode()
    header += b" " * ((-(len(MAGIC) + 8 + len(header))) % 64)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(header)))
        f.write(header)
        for e, a in zip(entries, arrays):
            f.write(a.tobytes())
            f.write(b"\0" * ((-a.nbytes) % 64))
    return os.path.getsize(path)


def load_blob(path: str) -> dict:
    """-> {prefix: {name: read-only memory-mapped fp32 array}}."""
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError(f"{path}: not a b200tts weight blob")
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode())
    base = len(MAGIC) + 8 + hlen
    size = os.path.getsize(path)
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    parts = {}
    for e in header["tensors"]:
        lo, hi = base + e["offset"], base + e["offset"] + e["nbytes"]
        if hi > size or e["nbytes"] != 4 * int(np.prod(e["shape"], dtype=np.int64)):
            raise ValueError(f"{path}: tensor '{e['name']}' is truncated or inconsistent")
        prefix, name = e["name"].split(".", 1)
        parts.setdefault(prefix, {})[name] = mm[lo:hi].view("<f4").reshape(e["shape"])
    return parts


def load_blob_into(engine, path: str):
    """Upload every part of a blob (capi.Engine.load_state per prefix). -> list of prefixes loaded."""
    parts = load_blob(path)
    for prefix, tensors in parts.items():
        engine.load_state(prefix, tensors)
    return list(parts)

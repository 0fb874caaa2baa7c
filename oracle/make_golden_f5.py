"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Build-container only.

F5-TTS golden vectors from the REFERENCE's own F5Preprocess / F5Transformer / F5Decode (oracle/ref_harness.py):
a small utterance (16384 samples -> 65 reference frames, 20 text ids, N = 130) through graph A, all 31 NFE steps of
graph B with the host loop of F5-TTS-ONNX-Inference.py:290-304, and graph C."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402,F401
from b200tts import config, synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DIT_SEED, VOCOS_SEED, INPUT_SEED = 4321, 2468, 1
AUDIO_LEN, N_TEXT = 16384, 20


def main():
    cfg = config.F5
    dsd, vsd = synth.f5_dit_state(DIT_SEED), synth.vocos_state(VOCOS_SEED)
    pre, trans, dec = ref_harness.build_f5(dsd, vsd, cfg)
    audio, text_ids, maxd, noise = synth.f5_inputs(INPUT_SEED, audio_len=AUDIO_LEN, n_text=N_TEXT)
    out = {"dit_seed": np.int64(DIT_SEED), "vocos_seed": np.int64(VOCOS_SEED), "input_seed": np.int64(INPUT_SEED),
           "audio_len": np.int64(AUDIO_LEN), "n_text": np.int64(N_TEXT)}
    with torch.inference_mode():
        a = pre(torch.from_numpy(audio), torch.from_numpy(text_ids), torch.from_numpy(maxd))
        out["cat_mel_text"] = a[5].numpy()
        out["cat_mel_text_drop"] = a[6].numpy()
        out["ref_signal_len"] = np.int64(a[7])
        out["rope_cos_row"] = a[1][0, 0].numpy()           # (N, 64); the graph output is this row repeated (2,16,.,.)
        out["rope_sin_row"] = a[2][0, 0].numpy()
        out["time_expand"] = trans.time_expand[0].numpy()
        out["delta_t"] = trans.delta_t.numpy()
        x = torch.from_numpy(noise).clone()
        ts = torch.tensor([0], dtype=torch.int32)
        for step in range(cfg.nfe - 1):
            x, ts = trans(x, a[1], a[2], a[3], a[4], a[5], a[6], ts)
            if step in (0, 1, 7, 30):
                out[f"noise_after_{step + 1}"] = x.numpy().copy()
        assert int(ts) == cfg.nfe - 1
        pcm = dec(x, torch.tensor(a[7]))
        out["pcm"] = pcm.numpy()
        # a decode-only vector on random mel (SURVEY s4 known answer shape)
        rng = np.random.default_rng(77)
        mel = rng.standard_normal((1, 40, cfg.n_mels), dtype=np.float32)
        out["decode_in"] = mel
        out["decode_pcm"] = dec(torch.from_numpy(mel), torch.tensor(12)).numpy()
    np.savez_compressed(os.path.join(GOLD, "f5_ref.npz"), **out)
    print("f5_ref.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})
    print("pcm rms", float(np.sqrt(np.mean((out["pcm"].astype(np.float64) / 32768) ** 2))), "max", int(np.abs(out["pcm"]).max()))


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Build-container only.

tests/golden/indextts_ref.npz: the reference's own IndexTTS BigVGAN module + gpt.final_norm, wrapped as IndexTTS_F
(oracle/ref_harness.py::build_indextts_f, sources imported from /root/reference/IndexTTS/modeling_modified), run on seeded
synthetic weights and inputs (text-to-speech-tts-onnx_b200/synth.py).

    python -m oracle.make_golden_indextts
"""
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


def main():
    cfg = config.INDEXTTS_VOCODER
    sd = synth.ivgan_state(777)
    ref = ref_harness.build_indextts_f(sd, cfg)
    out = {"weights_seed": np.int64(777)}
    for name, (seed, rows) in {"a": (5, 6), "b": (6, 11)}.items():
        conds, cond_layer, hidden = synth.ivgan_inputs(seed, rows)
        with torch.inference_mode():
            pcm = ref(*[torch.from_numpy(c) for c in conds], torch.from_numpy(cond_layer), torch.from_numpy(hidden))
        out[f"{name}_seed"] = np.int64(seed)
        out[f"{name}_rows"] = np.int64(rows)
        out[f"{name}_pcm"] = pcm.numpy()
    np.savez_compressed(os.path.join(GOLD, "indextts_ref.npz"), **out)
    print("indextts_ref.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()

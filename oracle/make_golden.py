"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Build-container only.

Generates tests/golden/*.npz by running the REFERENCE's own modules (oracle/ref_harness.py, imported from
/root/reference) on seeded synthetic checkpoints/inputs (text-to-speech-tts-onnx_b200/synth.py):

    python -m oracle.make_golden [bigvgan] [f5]

The fixtures pin the CPU restatement (tests/test_oracle_*.py, CPU) and the CUDA engine (tests/test_gpu_*.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402
from b200tts import config, synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def make_bigvgan():
    cfg = config.BIGVGAN
    sd = synth.bigvgan_state(1234)
    ref = ref_harness.build_bigvgan(sd, cfg)
    out = {"weights_seed": np.int64(1234)}
    for name, (seed, B, T) in {"a": (11, 1, 8), "b": (12, 2, 24)}.items():
        mel = synth.bigvgan_mel(seed, B, T)
        with torch.inference_mode():
            pcm = torch.cat([ref(torch.from_numpy(mel[i:i + 1])) for i in range(B)], 0)   # the reference graph is batch-1
        out[f"{name}_mel_seed"] = np.int64(seed)
        out[f"{name}_mel"] = mel
        out[f"{name}_pcm"] = pcm.numpy()
    # single-op vectors from the reference's own Activation1d (stage table and the index -1 "post" table)
    model = ref.bigvgan
    rng = np.random.default_rng(5)
    x = (2.0 * rng.standard_normal((2, 24, 53), dtype=np.float32))
    act = model.activation_post
    def run_act(idx):       # the reference's pad tables are batch-1 (bigvgan.py:359-382)
        with torch.inference_mode():
            return torch.cat([act(torch.from_numpy(x[i:i + 1]), model.x_shape[-1], model.up_filter_pad[idx],
                                  model.up_pad_zeros[idx], model.down_filter_pad[idx], model.down_pad_zeros_L[idx],
                                  model.down_pad_zeros_R[idx]) for i in range(x.shape[0])], 0)
    y_stage, y_post = run_act(5), run_act(-1)
    out["act_x"] = x
    out["act_alpha"] = sd["activation_post.act.alpha"]
    out["act_beta"] = sd["activation_post.act.beta"]
    out["act_filter"] = act.upsample.filter.reshape(-1).numpy()
    out["act_y_stage"] = y_stage.numpy()
    out["act_y_post"] = y_post.numpy()
    np.savez_compressed(os.path.join(GOLD, "bigvgan_ref.npz"), **out)
    print("bigvgan_ref.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    what = sys.argv[1:] or ["bigvgan", "f5"]
    os.makedirs(GOLD, exist_ok=True)
    if "bigvgan" in what:
        make_bigvgan()
    if "f5" in what:
        from oracle import make_golden_f5
        make_golden_f5.main()

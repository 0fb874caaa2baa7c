"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Build-container only.

Golden vectors at the BASELINE.json sizes, made by the REFERENCE's own modules (oracle/ref_harness.py, loaded from
/root/reference) on the seeded synthetic checkpoints / inputs the benchmark uses:

  * BigVGAN  configs[0]/[1]: mel (1,100,512) = synth.bigvgan_mel(100, 1, 512)  -> generated_wav int16 (1,1,131102)
  * F5-TTS   configs[2]    : synth.f5_inputs(1, 144000, 150) (6 s reference, 150 text ids, N = 1126) through graph A,
                             all 31 NFE steps of graph B with the host loop of F5-TTS-ONNX-Inference.py:290-304, graph C
                             -> denoised mel (1,1126,100) fp32 after steps 1 and 31, output_audio int16 (1,1,143872)

    python -m oracle.make_golden_fullsize        (about 3 minutes of CPU)

Inputs are regenerated from their seeds by the tests; only outputs are stored (tests/golden/fullsize_ref.npz, < 2 MB)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402,F401
from b200tts import config, synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
VGAN_SEED, VGAN_MEL_SEED, VGAN_T = 1234, 100, 512
DIT_SEED, VOCOS_SEED, INPUT_SEED, AUDIO_LEN, N_TEXT = 4321, 2468, 1, 144000, 150


def main():
    out = {"vgan_seed": np.int64(VGAN_SEED), "vgan_mel_seed": np.int64(VGAN_MEL_SEED), "vgan_T": np.int64(VGAN_T),
           "dit_seed": np.int64(DIT_SEED), "vocos_seed": np.int64(VOCOS_SEED), "input_seed": np.int64(INPUT_SEED),
           "audio_len": np.int64(AUDIO_LEN), "n_text": np.int64(N_TEXT)}
    t0 = time.time()
    ref = ref_harness.build_bigvgan(synth.bigvgan_state(VGAN_SEED), config.BIGVGAN)
    mel = synth.bigvgan_mel(VGAN_MEL_SEED, 1, VGAN_T)
    with torch.inference_mode():
        out["vgan_pcm"] = ref(torch.from_numpy(mel)).numpy()
    print("bigvgan", out["vgan_pcm"].shape, f"{time.time() - t0:.1f}s", flush=True)

    cfg = config.F5
    pre, trans, dec = ref_harness.build_f5(synth.f5_dit_state(DIT_SEED), synth.vocos_state(VOCOS_SEED), cfg)
    audio, text_ids, maxd, noise = synth.f5_inputs(INPUT_SEED, audio_len=AUDIO_LEN, n_text=N_TEXT)
    with torch.inference_mode():
        a = pre(torch.from_numpy(audio), torch.from_numpy(text_ids), torch.from_numpy(maxd))
        out["f5_ref_signal_len"] = np.int64(a[7])
        out["f5_cat_mel_text_sum"] = np.float64(a[5].double().sum())          # cheap pin of graph A at this size
        x = torch.from_numpy(noise).clone()
        ts = torch.tensor([0], dtype=torch.int32)
        for step in range(cfg.nfe - 1):
            x, ts = trans(x, a[1], a[2], a[3], a[4], a[5], a[6], ts)
            if step == 0:
                out["f5_mel_after_1"] = x.numpy().copy()
            print("step", step, f"{time.time() - t0:.1f}s", flush=True)
        out["f5_mel"] = x.numpy().copy()
        out["f5_pcm"] = dec(x, torch.tensor(a[7])).numpy()
    np.savez_compressed(os.path.join(GOLD, "fullsize_ref.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})
    print("f5 pcm rms", float(np.sqrt(np.mean((out["f5_pcm"].astype(np.float64) / 32768) ** 2))), "max", int(np.abs(out["f5_pcm"]).max()))


if __name__ == "__main__":
    main()

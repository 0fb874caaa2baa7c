"""CPU: the BigVGAN restatement (oracle/bigvgan_ref.py) against the vectors produced by the reference's own
modules (tests/golden/bigvgan_ref.npz, made by oracle/make_golden.py), plus the known answers of SURVEY.md s4."""
import numpy as np
import torch

import b200tts  # noqa: F401
from b200tts import config, synth, weights
from oracle import bigvgan_ref as R

CFG = config.BIGVGAN


def test_filter_known_answer():
    f = R.aa_filter().numpy()
    want = [0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971]
    np.testing.assert_allclose(f[:6], want, atol=2e-9)
    np.testing.assert_allclose(f, f[::-1], atol=1e-9)            # up- and down-filter are the same symmetric FIR
    np.testing.assert_array_equal(f, weights.kaiser_sinc_filter1d(0.25, 0.3, 12))   # product-side construction


def test_filter_matches_reference_buffer(golden_bigvgan):
    np.testing.assert_array_equal(R.aa_filter().numpy(), golden_bigvgan["act_filter"])


def test_activation_matches_reference(golden_bigvgan):
    g = golden_bigvgan
    x, a, b = (torch.from_numpy(g[k]) for k in ("act_x", "act_alpha", "act_beta"))
    y = R.activation1d(x, a, b, R.aa_filter())
    np.testing.assert_allclose(y.numpy(), g["act_y_stage"], rtol=0, atol=1e-6)
    y = R.activation1d(x, a, b, R.aa_filter(), 15, 15, 15)
    assert y.shape[-1] == x.shape[-1] + 30                        # quirk q2: +30 samples
    np.testing.assert_allclose(y.numpy(), g["act_y_post"], rtol=0, atol=1e-6)


def test_activation_zero_pad_not_replicate():
    """quirk q1: a constant input does NOT stay constant at the edges (zero pad by concat)."""
    x = torch.ones(1, 1, 40)
    y = R.activation1d(x, torch.zeros(1), torch.zeros(1), R.aa_filter())
    assert abs(float(y[0, 0, 20]) - (1 + np.sin(1.0) ** 2)) < 1e-3
    assert abs(float(y[0, 0, 0]) - float(y[0, 0, 20])) > 0.1


def test_bigvgan_pcm_bit_exact_vs_reference(golden_bigvgan):
    g = golden_bigvgan
    sd = synth.bigvgan_state(int(g["weights_seed"]))
    for tag in ("a", "b"):
        mel = g[f"{tag}_mel"]
        np.testing.assert_array_equal(mel, synth.bigvgan_mel(int(g[f"{tag}_mel_seed"]), *mel.shape[::2]))
        # the reference graph is batch-1 (its pad tables are (1, C, pad)): item by item the restatement is bit-exact
        pcm = np.concatenate([R.bigvgan_pcm(mel[i:i + 1], sd, CFG).numpy() for i in range(mel.shape[0])], 0)
        assert pcm.dtype == np.int16 and pcm.shape == (mel.shape[0], 1, 256 * mel.shape[2] + 30)
        np.testing.assert_array_equal(pcm, g[f"{tag}_pcm"])
        # batched, the CPU conv kernels pick another blocking: fp32 re-association moves a few samples by 1 LSB
        pcm_b = R.bigvgan_pcm(mel, sd, CFG).numpy()
        assert np.abs(pcm_b.astype(np.int32) - g[f"{tag}_pcm"]).max() <= 1


def test_output_length_and_dtype():
    assert CFG.hop == 256
    assert CFG.out_samples(32) == 8222 and CFG.out_samples(512) == 131102
    n = sum(v.size for v in synth.bigvgan_state(1).values())
    assert abs(n / 1e6 - 112.4) < 0.1                              # 112.4 M parameters

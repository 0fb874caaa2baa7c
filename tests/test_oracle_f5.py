"""CPU: the F5 restatement (oracle/f5_ref.py) against vectors produced by the reference's own F5Preprocess /
F5Transformer / F5Decode (tests/golden/f5_ref.npz, oracle/make_golden_f5.py) and the known answers of SURVEY.md s4."""
import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import config, synth
from oracle import f5_ref as R

CFG = config.F5
GOLD = None


@pytest.fixture(scope="module")
def g():
    import os
    from conftest import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, "f5_ref.npz")))


@pytest.fixture(scope="module")
def dit_sd(g):
    return R.prescale_qk(synth.f5_dit_state(int(g["dit_seed"])), CFG)


def test_time_grid_known_answers(g, dit_sd):
    delta_t, time_expand, t = R.time_tables(dit_sd, CFG)
    np.testing.assert_allclose(t[:4].numpy(), [0, 0.00128353, 0.00513071, 0.01153165], atol=2e-7)
    assert abs(float(t[31]) - 1.0) < 1e-6 and delta_t.shape == (31,) and abs(float(delta_t.sum()) - 1.0) < 1e-6
    np.testing.assert_array_equal(time_expand.numpy(), g["time_expand"])
    np.testing.assert_array_equal(delta_t.numpy(), g["delta_t"])


def test_rope_fp16_rounding(g):
    cos, sin = R.rope_tables(CFG)
    np.testing.assert_array_equal(cos[:130].numpy(), g["rope_cos_row"])
    np.testing.assert_array_equal(sin[:130].numpy(), g["rope_sin_row"])
    hd = CFG.head_dim
    exact = torch.outer(torch.arange(CFG.max_frames, dtype=torch.float32),
                        1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))).repeat_interleave(2, dim=-1).cos()
    d = float((cos - exact).abs().max())
    assert 1e-4 < d <= 2.45e-4                      # quirk q5: the tables went through fp16


def test_mel_fbank_known_answers():
    fb = R.mel_fbank(CFG)[0].T                      # (513, 100)
    assert fb.shape == (513, 100) and abs(float(fb.max()) - 0.99903) < 1e-4
    np.testing.assert_allclose(fb.sum(0)[:3].numpy(), [0.8541, 0.8899, 0.9237], atol=1e-3)


def test_stft_istft_known_answers():
    torch.manual_seed(0)
    x = torch.randn(1, 1, 144000)
    re, im = R.stft_B(x)
    assert re.shape == (1, 513, 563)
    want = torch.stft(x[0, 0], 1024, 256, 1024, torch.hann_window(1024), center=True, pad_mode="reflect", return_complex=True)
    assert float((torch.complex(re[0], im[0]) - want).abs().max()) < 1e-2      # fp32-angle basis, magnitude ~70 (q8)
    mag, ph = torch.sqrt(re * re + im * im), torch.atan2(im, re)
    y = R.istft_A(mag, ph, R.istft_tables())
    assert y.shape == (1, 1, 256 * 562)
    err = (y[0, 0] - x[0, 0, : y.shape[-1]]).abs()
    assert float(err[1024:-1024].max()) < 1e-3          # atan2 round trip; interior is clean
    assert float(err[-512:].max()) > 0.05                                      # quirk q9: under-normalised tail


def test_preprocess_matches_reference(g, dit_sd):
    audio, text_ids, maxd, noise = synth.f5_inputs(int(g["input_seed"]), int(g["audio_len"]), int(g["n_text"]))
    out = R.f5_preprocess(audio, text_ids, maxd, dit_sd, CFG, noise=noise)
    assert out[7] == int(g["ref_signal_len"]) == int(g["audio_len"]) // 256 + 1
    np.testing.assert_array_equal(out[5].numpy(), g["cat_mel_text"])
    np.testing.assert_array_equal(out[6].numpy(), g["cat_mel_text_drop"])
    assert out[1].shape == (2, 16, 130, 64) and out[3].shape == (2, 16, 64, 130)
    assert float(out[6][0, :, :100].abs().max()) == 0.0           # cond_drop mel half is zeros
    # filler rows (beyond the text) are exactly zero in the text half (masked_fill), for text AND text_drop (q11)
    assert float(out[5][0, int(g["n_text"]):, 100:].abs().max()) == 0.0
    assert float(out[6][0, int(g["n_text"]):, 100:].abs().max()) == 0.0


def test_transformer_steps_match_reference(g, dit_sd):
    _, _, _, noise = synth.f5_inputs(int(g["input_seed"]), int(g["audio_len"]), int(g["n_text"]))
    tables = R.time_tables(dit_sd, CFG)
    x, ts = torch.from_numpy(noise), 0
    cond, drop = torch.from_numpy(g["cat_mel_text"]), torch.from_numpy(g["cat_mel_text_drop"])
    cos, sin = torch.from_numpy(g["rope_cos_row"]), torch.from_numpy(g["rope_sin_row"])
    for step in range(2):
        x, ts = R.f5_transformer_step(dit_sd, x, cond, drop, ts, tables, CFG, cos, sin)
        np.testing.assert_allclose(x.numpy(), g[f"noise_after_{step + 1}"], rtol=0, atol=1e-6)
    assert ts == 2


def test_decode_matches_reference(g):
    fsd = R.fold_vocos(synth.vocos_state(int(g["vocos_seed"])), CFG)
    pcm = R.f5_decode(g["decode_in"], 12, fsd, CFG).numpy()
    assert pcm.dtype == np.int16 and pcm.shape == g["decode_pcm"].shape == (1, 1, 256 * (40 - 12 - 1))
    assert np.abs(pcm.astype(np.int32) - g["decode_pcm"]).max() <= 1
    pcm = R.f5_decode(g["noise_after_31"], int(g["ref_signal_len"]), fsd, CFG).numpy()
    assert np.abs(pcm.astype(np.int32) - g["pcm"]).max() <= 1
    # quirk q12: ORT may swap exact GELU for the tanh approximation at run time; the effect is bounded
    mag_e, _ = R.vocos_decode(fsd, torch.from_numpy(g["decode_in"]).transpose(1, 2), CFG, gelu_tanh=False)
    mag_t, _ = R.vocos_decode(fsd, torch.from_numpy(g["decode_in"]).transpose(1, 2), CFG, gelu_tanh=True)
    assert float(((mag_e - mag_t).abs() / (mag_e.abs() + 1e-3)).max()) < 0.05

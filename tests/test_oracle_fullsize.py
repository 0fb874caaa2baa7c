"""CPU: the oracle restatements at BASELINE.json sizes against tests/golden/fullsize_ref.npz, which the reference's own
modules produced (oracle/make_golden_fullsize.py). BigVGAN (1,100,512) in full; F5 N = 1126: graph A and the first of the 31
Euler steps (all 31 take two minutes of CPU -- the GPU suite checks the engine against the stored 31-step mel / PCM)."""
import os

import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import config, synth
from conftest import GOLDEN
from oracle import bigvgan_ref, f5_ref


@pytest.fixture(scope="module")
def gf():
    return dict(np.load(os.path.join(GOLDEN, "fullsize_ref.npz")))


def test_bigvgan_oracle_config1_bit_exact(gf):
    mel = synth.bigvgan_mel(int(gf["vgan_mel_seed"]), 1, int(gf["vgan_T"]))
    pcm = bigvgan_ref.bigvgan_pcm(mel, synth.bigvgan_state(int(gf["vgan_seed"])), config.BIGVGAN).numpy()
    assert pcm.shape == gf["vgan_pcm"].shape == (1, 1, 131102)
    d = np.abs(pcm.astype(np.int32) - gf["vgan_pcm"].astype(np.int32))
    assert d.max() <= 1 and (d == 0).mean() > 0.999          # same modules' arithmetic, restated; fp32 summation order of conv


def test_f5_oracle_config3_first_step(gf):
    cfg = config.F5
    dsd = synth.f5_dit_state(int(gf["dit_seed"]))
    audio, text_ids, maxd, noise = synth.f5_inputs(int(gf["input_seed"]), int(gf["audio_len"]), int(gf["n_text"]))
    with torch.inference_mode():
        sd = f5_ref.prescale_qk(dsd, cfg)
        tables = f5_ref.time_tables(sd, cfg)
        x, cq, sq, _, _, cond, cond_drop, ref_len = f5_ref.f5_preprocess(audio, text_ids, maxd, sd, cfg, noise)
        assert int(ref_len) == int(gf["f5_ref_signal_len"]) == 563 and x.shape == (1, 1126, 100)
        assert abs(float(cond.double().sum()) - float(gf["f5_cat_mel_text_sum"])) <= 1e-6 * abs(float(gf["f5_cat_mel_text_sum"])) + 1e-2
        x1, ts = f5_ref.f5_transformer_step(sd, x, cond, cond_drop, 0, tables, cfg, cq[0, 0], sq[0, 0])
    assert int(ts) == 1
    np.testing.assert_allclose(x1.numpy(), gf["f5_mel_after_1"], rtol=0, atol=2e-5)

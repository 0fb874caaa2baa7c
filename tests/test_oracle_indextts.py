"""CPU: the IndexTTS_F restatement (oracle/indextts_ref.py) against the vectors produced by the reference's own IndexTTS BigVGAN
module wrapped as IndexTTS_F (tests/golden/indextts_ref.npz, made by oracle/make_golden_indextts.py)."""
import os

import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import config, synth, weights
from conftest import GOLDEN
from oracle import indextts_ref as R

CFG = config.INDEXTTS_VOCODER


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "indextts_ref.npz")))


@pytest.mark.parametrize("case", ["a", "b"])
def test_pcm_bit_exact_vs_reference(g, case):
    sd = synth.ivgan_state(int(g["weights_seed"]))
    conds, cond_layer, hidden = synth.ivgan_inputs(int(g[f"{case}_seed"]), int(g[f"{case}_rows"]))
    pcm = R.indextts_f_pcm(hidden, conds, cond_layer, sd, CFG).numpy()
    assert pcm.dtype == np.int16 and pcm.shape == g[f"{case}_pcm"].shape
    np.testing.assert_array_equal(pcm, g[f"{case}_pcm"])


def test_output_length_drops_two_rows_and_adds_30():
    # x1024 ladder, the last two latent rows are dropped (Export_IndexTTS.py:301), the 15-pad post tables add 30 samples
    assert CFG.hop == 1024 and CFG.out_samples(142 - 2) == 140 * 1024 + 30
    sd = synth.ivgan_state(777)
    conds, cond_layer, hidden = synth.ivgan_inputs(1, 3)
    assert tuple(R.indextts_f_pcm(hidden, conds, cond_layer, sd, CFG).shape) == (1, 1, 1024 + 30)


def test_conditioning_and_post_bias_matter():
    sd = synth.ivgan_state(777)
    conds, cond_layer, hidden = synth.ivgan_inputs(2, 4)
    base = R.indextts_f_forward(hidden, conds, cond_layer, sd, CFG)
    zero = [np.zeros_like(c) for c in conds]
    assert float((R.indextts_f_forward(hidden, zero, cond_layer, sd, CFG) - base).abs().max()) > 1e-4
    sd2 = dict(sd); sd2["conv_post.bias"] = sd["conv_post.bias"] + 0.05
    assert float((R.indextts_f_forward(hidden, conds, cond_layer, sd2, CFG) - base).abs().max()) > 1e-3


def test_engine_tensor_names():
    t = weights.ivgan_engine_tensors(synth.ivgan_state(777), CFG)
    assert t["upsample_rates"].tolist() == [4, 4, 4, 4, 2, 2] and t["aa_filter"].shape == (12,)
    assert t["conv_pre.weight"].shape == (1536, 1280, 7) and t["ups.2.0.weight"].shape == (384, 192, 4)
    assert "final_norm.weight" in t and "conv_post.bias" in t

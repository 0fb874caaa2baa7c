"""CPU: the IndexTTS GPT-2 decode restatement (oracle/indextts_gpt_ref.py: graphs B-E + the host loop) against the vectors
produced by the reference's own IndexTTS_B/C/D/E classes around a Hugging Face GPT2Model (tests/golden/indextts_gpt_ref.npz,
made by oracle/make_golden_indextts_gpt.py)."""
import os

import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import config, synth
from conftest import GOLDEN
from oracle import indextts_gpt_ref as R

CFG = config.INDEXTTS_GPT_SMALL


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "indextts_gpt_ref.npz")))


@pytest.fixture(scope="module")
def run(g):
    sd = synth.igpt_state(int(g["seed_w"]), CFG)
    conds, text_ids = synth.igpt_inputs(int(g["seed_in"]), int(g["n_text"]), CFG)
    return sd, conds, text_ids, R.generate(conds, text_ids, sd, CFG, max_new=int(g["max_new"]))


def test_text_embedding_bit_exact(g, run):
    sd, conds, text_ids, _ = run
    t = R.text_embed(text_ids, sd, CFG).numpy()
    assert t.shape == (1, text_ids.shape[1] + 2, CFG.dim)         # start and stop ids are added inside graph B
    np.testing.assert_array_equal(t, g["text_hidden"])


def test_greedy_ids_and_hidden_states_vs_reference(g, run):
    ids, hidden, pen = run[3]
    np.testing.assert_array_equal(ids, g["ids"])                 # 40 greedy tokens, penalty window active after the 10th
    assert hidden.shape == g["hidden"].shape
    assert np.abs(hidden - g["hidden"]).max() <= 2e-5            # same torch ops in the same order; fp32 round-off only
    np.testing.assert_array_equal(pen, g["penalty"])


def test_penalty_window_releases_oldest(run):
    ids, _, pen = run[3]
    n = len(ids)
    held = set(np.nonzero(pen[0] != 1.0)[0].tolist())
    assert np.all(pen[0][list(held)] == np.float32(CFG.repeat_penalty))
    # every token of the last PENALITY_RANGE calls is still penalised; the very first one was released (unless repeated later)
    assert set(ids[-CFG.penalty_range:-1].tolist()) <= held
    assert n <= CFG.penalty_range or ids[0] in ids[1:].tolist() or ids[0] not in held


def test_mask_flag_is_additive_minus_128(run):
    """Export_IndexTTS.py:245,268: the 'mask' adds -128 above the diagonal (not -inf) and only when the int8 flag is 1."""
    sd, conds, text_ids, _ = run
    folded = [R.fold_layer(sd, i, CFG) for i in range(CFG.layers)]
    H, hd = CFG.heads, CFG.head_dim
    pk = [torch.zeros((H, hd, 0)) for _ in range(CFG.layers)]
    pv = [torch.zeros((H, 0, hd)) for _ in range(CFG.layers)]
    x = torch.from_numpy(conds)
    pen = torch.ones((1, CFG.mel_codes))
    a = R.step_e(folded, pk, pv, pen, x, 1, sd, CFG)
    b = R.step_e(folded, pk, pv, pen, x, 0, sd, CFG)
    assert not torch.allclose(a[2], b[2])                        # without the flag every row sees the whole block
    # causal prefill == feeding the rows one at a time with the flag off
    for r in range(x.shape[1]):
        pk, pv, last, _, _ = R.step_e(folded, pk, pv, pen, x[:, r:r + 1], 0, sd, CFG)
    assert torch.allclose(last, a[2], atol=2e-5)

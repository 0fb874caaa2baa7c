"""CPU: host logic of the onnxruntime-shaped IndexTTS sessions (session.py: graphs B, C, D, E, ResidentKV handles) driven by the
reference script's loop (Inference_IndexTTS_ONNX.py:719-781), with the device engine replaced by a stand-in that answers the
same capi calls from the oracle. Checks names, ordering, dtypes and the handle plumbing -- the arithmetic is the GPU tests' job."""
import os

import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import config, session as ort, synth
from conftest import GOLDEN
from oracle import indextts_gpt_ref as R

CFG = config.INDEXTTS_GPT_SMALL


class FakeEngine:
    def __init__(self, sd):
        self.sd = sd
        self.folded = [R.fold_layer(sd, i, CFG) for i in range(CFG.layers)]
        self.reset()
        self.loaded = []

    def reset(self):
        self.k = [torch.zeros((CFG.heads, 64, 0)) for _ in range(CFG.layers)]
        self.v = [torch.zeros((CFG.heads, 0, 64)) for _ in range(CFG.layers)]

    def load_state(self, prefix, tensors):
        self.loaded.append((prefix, sorted(tensors)))

    def indextts_gpt_build(self):
        pass

    def indextts_gpt_text_embed(self, ids):
        return R.text_embed(ids, self.sd, CFG).numpy()

    def indextts_gpt_mel_embed(self, mel_id, gen_len):
        h, nxt = R.mel_embed([[mel_id]], gen_len, self.sd)
        return h.numpy(), nxt

    def indextts_gpt_step(self, hidden, history_len, attention_mask, pen, precision=0):
        if history_len == 0:
            self.reset()
        assert history_len == self.k[0].shape[2], "history_len does not match the resident KV cache"
        self.k, self.v, last, mid, _ = R.step_e(self.folded, self.k, self.v, torch.from_numpy(np.asarray(pen)), torch.from_numpy(hidden),
                                                attention_mask, self.sd, CFG)
        return last.numpy(), mid.numpy().astype(np.int32), self.k[0].shape[2]

    def indextts_gpt_kv_read(self, layer):
        return self.k[layer].numpy(), self.v[layer].numpy()


@pytest.fixture()
def sessions(monkeypatch):
    g = dict(np.load(os.path.join(GOLDEN, "indextts_gpt_ref.npz")))
    sd = synth.igpt_state(int(g["seed_w"]), CFG)
    fake = FakeEngine(sd)
    monkeypatch.setattr(ort, "get_engine", lambda device=0: fake)
    ort.register_checkpoint("indextts_gpt", sd)
    ort.register_checkpoint("indextts_gpt_config", CFG)
    ort._igpt_ready.clear()
    s = {k: ort.InferenceSession(f"/x/IndexTTS_{k}.onnx", precision="fp32") for k in "BCDE"}
    yield g, fake, s
    ort._igpt_ready.clear()
    ort._checkpoints.pop("indextts_gpt", None)
    ort._checkpoints.pop("indextts_gpt_config", None)


def test_checkpoint_is_loaded_once_and_names_match_the_export(sessions):
    g, fake, s = sessions
    assert [p for p, _ in fake.loaded] == ["igpt"]                       # B, C and E share one model
    assert "meta" in fake.loaded[0][1] and "h.0.attn.c_attn.weight" in fake.loaded[0][1]
    L = CFG.layers
    names_in = [i.name for i in s["E"].get_inputs()]
    names_out = [o.name for o in s["E"].get_outputs()]
    assert names_in == [f"in_key_{i}" for i in range(L)] + [f"in_value_{i}" for i in range(L)] + [
        "history_len", "repeat_penality", "ids_len", "hidden_state", "attention_mask"]          # Export_IndexTTS.py:428-458
    assert names_out == [f"out_key_{i}" for i in range(L)] + [f"out_value_{i}" for i in range(L)] + [
        "kv_seq_len", "last_hidden_state", "max_logit_id"]
    assert [i.name for i in s["B"].get_inputs()] == ["text_ids"] and [o.name for o in s["C"].get_outputs()] == ["gpt_hidden_state", "next_kv_seq_len"]
    assert [o.name for o in s["D"].get_outputs()] == ["concat_hidden_state", "concat_len"]
    assert "float16" not in s["E"]._inputs_meta[0].type and s["E"]._inputs_meta[0].shape[:2] == [CFG.heads, 64]


def test_reference_loop_through_the_sessions(sessions):
    g, fake, s = sessions
    OV = ort.OrtValue
    conds, text_ids = synth.igpt_inputs(int(g["seed_in"]), int(g["n_text"]), CFG)
    in_E = [i.name for i in s["E"].get_inputs()]
    out_E = [o.name for o in s["E"].get_outputs()]
    L = CFG.layers
    text_h = s["B"].run_with_ort_values(["text_hidden_state"], {"text_ids": OV.ortvalue_from_numpy(text_ids)})[0]
    gpt_h, gen_len = s["C"].run_with_ort_values(None, {"gpt_ids": OV.ortvalue_from_numpy(np.array([[CFG.start_mel]], np.int32)),
                                                       "kv_seq_len": OV.ortvalue_from_numpy(np.array([0], np.int64))})
    gpt_h, concat_len = s["D"].run_with_ort_values(None, {"embed_x": OV.ortvalue_from_numpy(conds), "embed_y": text_h, "embed_z": gpt_h})
    np.testing.assert_array_equal(OV.numpy(gpt_h), g["concat_hidden"])
    assert OV.numpy(concat_len).dtype == np.int64 and int(OV.numpy(concat_len)[0]) == g["concat_hidden"].shape[1]
    feed = {n: OV.ortvalue_from_numpy(np.zeros((CFG.heads, 64, 0), np.float32)) for n in in_E[:L]}
    feed.update({n: OV.ortvalue_from_numpy(np.zeros((CFG.heads, 0, 64), np.float32)) for n in in_E[L:2 * L]})
    pen = np.ones((1, CFG.mel_codes), np.float32)
    feed.update({"history_len": OV.ortvalue_from_numpy(np.array([0], np.int64)), "repeat_penality": OV.ortvalue_from_numpy(pen),
                 "ids_len": concat_len, "attention_mask": OV.ortvalue_from_numpy(np.array([1], np.int8))})
    ids, hid, reset = [], [], 0
    for n in range(1, 15):
        feed["hidden_state"] = gpt_h
        outs = s["E"].run_with_ort_values(out_E, feed)
        assert all(isinstance(o, ort.OrtValue) for o in outs) and isinstance(outs[0], ort.ResidentKV)
        mid = OV.numpy(outs[-1])
        assert mid.dtype == np.int32 and mid.shape == (1, 1)
        ids.append(int(mid[0, 0]))
        hid.append(OV.numpy(outs[-2])[0])
        if n < 2:
            feed["attention_mask"] = OV.ortvalue_from_numpy(np.array([0], np.int8))
            feed["ids_len"] = OV.ortvalue_from_numpy(np.array([1], np.int64))
        for i in range(2 * L + 1):
            feed[in_E[i]] = outs[i]
        pen = OV.numpy(feed["repeat_penality"])
        pen[:, ids[-1]] = CFG.repeat_penalty
        if n > CFG.penalty_range and ids[reset] != ids[-1]:
            pen[:, ids[reset]] = 1.0
            reset += 1
        feed["repeat_penality"] = OV.ortvalue_from_numpy(pen)
        gpt_h, gen_len = s["C"].run_with_ort_values(None, {"gpt_ids": outs[-1], "kv_seq_len": gen_len})
    np.testing.assert_array_equal(np.asarray(ids, np.int32), g["ids"][:14])
    assert np.abs(np.stack(hid) - g["hidden"][:14]).max() <= 2e-5
    S = g["concat_hidden"].shape[1] + 13
    assert OV.numpy(outs[0]).shape == (CFG.heads, 64, S) and outs[L].shape() == [CFG.heads, S, 64]
    assert int(OV.numpy(outs[2 * L])[0]) == S and int(OV.numpy(gen_len)[0]) == 15
    # KV handles are views of the resident cache: one more call makes the older ones stale -- reading them or feeding
    # them back raises instead of silently showing / attending over the advanced cache; host arrays are refused for history > 0
    old = list(outs)
    feed["hidden_state"] = gpt_h
    outs = s["E"].run_with_ort_values(out_E, feed)
    assert outs[0].is_current() and not old[0].is_current()
    with pytest.raises(RuntimeError, match="stale KV handle"):
        OV.numpy(old[0])
    stale = dict(feed)
    for i in range(2 * L + 1):
        stale[in_E[i]] = outs[i]
    stale[in_E[0]] = old[0]
    with pytest.raises(ValueError, match="stale or foreign KV handle"):
        s["E"].run_with_ort_values(out_E, stale)
    stale[in_E[0]] = OV.ortvalue_from_numpy(np.zeros((CFG.heads, 64, S + 1), np.float32))
    with pytest.raises(ValueError, match="RESIDENT cache"):
        s["E"].run_with_ort_values(out_E, stale)


def test_missing_inputs_and_bad_shapes_raise(sessions):
    g, fake, s = sessions
    with pytest.raises(ValueError, match="missing inputs"):
        s["E"].run(None, {"hidden_state": np.zeros((1, 1, CFG.dim), np.float32)})
    feed = {i.name: np.zeros((CFG.heads, 64, 0), np.float32) for i in s["E"].get_inputs()}
    feed.update({"history_len": np.array([0]), "repeat_penality": np.ones((1, CFG.mel_codes), np.float32), "ids_len": np.array([2]),
                 "hidden_state": np.zeros((1, 3, CFG.dim), np.float32), "attention_mask": np.array([1], np.int8)})
    with pytest.raises(ValueError, match="ids_len"):
        s["E"].run(None, feed)

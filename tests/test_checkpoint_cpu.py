"""CPU: checkpoint transforms (EMA selection, weight-norm removal) against torch's own implementations, and the weight blob."""
import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import checkpoint, config, synth, weights


@pytest.mark.parametrize("style", ["legacy", "parametrize"])
def test_remove_weight_norm_matches_torch(style):
    torch.manual_seed(0)
    conv = torch.nn.Conv1d(6, 8, 5)
    tr = torch.nn.ConvTranspose1d(8, 4, 4, stride=2)
    if style == "legacy":
        m = torch.nn.Sequential(torch.nn.utils.weight_norm(conv), torch.nn.utils.weight_norm(tr))
    else:
        m = torch.nn.Sequential(torch.nn.utils.parametrizations.weight_norm(conv), torch.nn.utils.parametrizations.weight_norm(tr))
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.3 * torch.randn_like(p))             # g != ||v||, so the product is not just v
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    assert any(k.endswith(("weight_g", "original0")) for k in sd)
    if style == "legacy":                                   # the hook recomputes .weight only in forward: let torch fold it
        for i in range(2):
            torch.nn.utils.remove_weight_norm(m[i])
    want = {f"{i}.weight": m[i].weight.detach().numpy().copy() for i in range(2)}
    got = checkpoint.remove_weight_norm(sd)
    assert sorted(got) == ["0.bias", "0.weight", "1.bias", "1.weight"]
    for k, w in want.items():
        np.testing.assert_allclose(got[k], w, rtol=1e-6, atol=1e-7)


def test_incomplete_weight_norm_pair_is_an_error():
    with pytest.raises(ValueError, match="incomplete"):
        checkpoint.remove_weight_norm({"conv.weight_g": np.ones((2, 1, 1), np.float32)})


def test_strip_ema_and_dit_selection():
    dit = {"time_embed.time_mlp.0.weight": np.ones((2, 2), np.float32), "proj_out.bias": np.zeros(3, np.float32)}
    ck = {"ema_model." + "transformer." + k: v for k, v in dit.items()}
    ck.update({"initted": np.array(1.0), "step": np.array(7), "ema_model.mel_spec.mel_stft.mel_scale.fb": np.zeros((3, 3), np.float32),
               "ema_model.mel_spec.mel_stft.spectrogram.window": np.zeros(4, np.float32)})
    out = checkpoint.f5_dit_from_checkpoint(ck)
    assert sorted(out) == sorted(dit)
    assert checkpoint.f5_dit_from_checkpoint({"ema_model_state_dict": ck}).keys() == out.keys()


def test_indextts_gpt_key_mapping_round_trip():
    cfg = config.INDEXTTS_GPT_SMALL
    sd = synth.igpt_state(1, cfg)
    up = {}
    for k, v in sd.items():
        up[("gpt." + k) if k.startswith(("h.", "ln_f.")) else k] = v
    up["gpt.h.0.attn.bias"] = np.ones((1, 1, 4, 4), np.float32)           # GPT2Attention's causal-mask buffer
    up["conditioning_encoder.x"] = np.zeros(3, np.float32)
    got = checkpoint.indextts_gpt_from_checkpoint({"model": up})
    assert sorted(got) == sorted(sd)


def test_blob_round_trip(tmp_path):
    parts = {"bigvgan": weights.bigvgan_engine_tensors({k: v for k, v in list(synth.bigvgan_state(3).items())[:12]}),
             "igpt": {"meta": np.arange(8, dtype=np.float32), "h.0.ln_1.weight": np.random.default_rng(0).standard_normal(512).astype(np.float32)}}
    p = str(tmp_path / "w.b200tts")
    n = checkpoint.save_blob(p, parts)
    assert n % 64 == 0
    back = checkpoint.load_blob(p)
    assert sorted(back) == sorted(parts)
    for prefix in parts:
        assert sorted(back[prefix]) == sorted(parts[prefix])
        for k, v in parts[prefix].items():
            a = back[prefix][k]
            assert a.dtype == np.float32 and a.shape == np.asarray(v).shape
            np.testing.assert_array_equal(a, np.asarray(v, dtype=np.float32))
    with open(p, "r+b") as f:
        f.truncate(n - 128)
    with pytest.raises(ValueError, match="truncated"):
        checkpoint.load_blob(p)
    with open(p, "r+b") as f:
        f.write(b"XXXXXXXX")
    with pytest.raises(ValueError, match="not a b200tts"):
        checkpoint.load_blob(p)

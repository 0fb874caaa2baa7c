"""GPU: a weight-normed 'upstream' BigVGAN checkpoint -> converter -> blob -> engine gives the same PCM as loading the state dict."""
import numpy as np
import pytest

import b200tts  # noqa: F401
from b200tts import capi, checkpoint, synth, weights

pytestmark = pytest.mark.gpu


def test_blob_loaded_bigvgan_matches_state_dict(engine, tmp_path):
    sd = synth.bigvgan_state(1234)
    rng = np.random.default_rng(5)
    up = {}
    for k, v in sd.items():                       # re-express every conv weight as a weight-norm pair, as the upstream file stores it
        if k.endswith(".weight") and v.ndim == 3:
            g = np.sqrt((v.reshape(v.shape[0], -1).astype(np.float64) ** 2).sum(1)).astype(np.float32).reshape(-1, 1, 1)
            s = rng.uniform(0.5, 2.0, size=g.shape).astype(np.float32)
            up[k[:-7] + ".weight_g"], up[k[:-7] + ".weight_v"] = g, v * s
        else:
            up[k] = v
    conv = checkpoint.bigvgan_from_checkpoint({"generator": up})
    assert sorted(conv) == sorted(sd)
    p = str(tmp_path / "vgan.b200tts")
    checkpoint.save_blob(p, {"bigvgan": weights.bigvgan_engine_tensors(conv)})
    mel = synth.bigvgan_mel(9, 1, 24)
    engine.load_state("bigvgan", weights.bigvgan_engine_tensors(sd))
    engine.bigvgan_build()
    want = engine.bigvgan_run(mel, precision=capi.F32)
    assert engine.load_blob(p) == ["bigvgan"]
    engine.bigvgan_build()
    got = engine.bigvgan_run(mel, precision=capi.F32)
    assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1      # g*v/||v|| reproduces the weights to fp32 round-off


def test_session_runs_from_a_blob_loaded_engine(tmp_path):
    """weights uploaded from an engine blob (Engine.load_blob) drive the onnxruntime-shaped sessions directly -- no
    register_checkpoint, no second conversion / upload."""
    from b200tts import session as ort
    sd = synth.bigvgan_state(77)
    p = str(tmp_path / "vgan.b200tts")
    checkpoint.save_blob(p, {"bigvgan": weights.bigvgan_engine_tensors(sd)})
    eng = capi.Engine(0)
    assert not eng.has_state("bigvgan")
    assert eng.load_blob(p) == ["bigvgan"] and eng.has_state("bigvgan")
    saved_engine, saved_ckpt = ort._engines.get(0), ort._checkpoints.pop("bigvgan", None)
    try:
        ort._engines[0] = eng
        sess = ort.InferenceSession("BigVGAN.onnx", precision="fp32")
        mel = synth.bigvgan_mel(3, 1, 16)
        got = sess.run(["generated_wav"], {"mel_features": mel})[0]
        eng.load_state("bigvgan", weights.bigvgan_engine_tensors(sd))
        eng.bigvgan_build()
        want = eng.bigvgan_run(mel, precision=capi.F32)
        np.testing.assert_array_equal(got, want)
    finally:
        if saved_engine is not None:
            ort._engines[0] = saved_engine
        else:
            ort._engines.pop(0, None)
        if saved_ckpt is not None:
            ort._checkpoints["bigvgan"] = saved_ckpt

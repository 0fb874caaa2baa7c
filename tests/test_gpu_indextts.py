"""GPU parity for the IndexTTS_F vocoder (SURVEY.md 8a row c1, config 5's vocoder half): CUDA engine through the C ABI vs the
vectors made by the reference's own IndexTTS BigVGAN module (tests/golden/indextts_ref.npz) and vs the oracle at another size.

Stated tolerances: fp32 engine PCM within 2 LSB of the reference's int16; bf16 engine PCM SNR >= 30 dB vs the fp32 oracle."""
import os

import numpy as np
import pytest

import b200tts  # noqa: F401
from b200tts import capi, config, synth, weights
from conftest import GOLDEN, snr_db
from oracle import indextts_ref as R

pytestmark = pytest.mark.gpu
CFG = config.INDEXTTS_VOCODER


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "indextts_ref.npz")))


@pytest.fixture(scope="module")
def ivgan(engine, g):
    engine.load_state("ivgan", weights.ivgan_engine_tensors(synth.ivgan_state(int(g["weights_seed"])), CFG))
    engine.indextts_vocoder_build()
    return engine


@pytest.mark.parametrize("case", ["a", "b"])
def test_f32_vs_reference_golden(ivgan, g, case):
    conds, cond_layer, hidden = synth.ivgan_inputs(int(g[f"{case}_seed"]), int(g[f"{case}_rows"]))
    pcm = ivgan.indextts_vocoder_run(hidden, conds, cond_layer, precision=capi.F32, hop=CFG.hop)
    want = g[f"{case}_pcm"]
    assert pcm.dtype == np.int16 and pcm.shape == want.shape
    assert np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 2


def test_bf16_vs_oracle(ivgan, g):
    sd = synth.ivgan_state(int(g["weights_seed"]))
    conds, cond_layer, hidden = synth.ivgan_inputs(21, 34)                  # 32 latent rows -> 32 798 samples
    want, wantf = R.indextts_f_pcm(hidden, conds, cond_layer, sd, CFG, return_float=True)
    pcm, wave = ivgan.indextts_vocoder_run(hidden, conds, cond_layer, precision=capi.BF16, return_wave=True, hop=CFG.hop)
    assert pcm.shape == tuple(want.shape)
    assert snr_db(wantf.numpy(), wave) > 30.0
    pcm32 = ivgan.indextts_vocoder_run(hidden, conds, cond_layer, precision=capi.F32, hop=CFG.hop)
    assert np.abs(pcm32.astype(np.int32) - want.numpy().astype(np.int32)).max() <= 2


def test_minimum_length_and_errors(ivgan):
    conds, cond_layer, hidden = synth.ivgan_inputs(3, 3)                    # one latent row survives
    assert ivgan.indextts_vocoder_run(hidden, conds, cond_layer, precision=capi.F32, hop=CFG.hop).shape == (1, 1, 1024 + 30)
    with pytest.raises((RuntimeError, AssertionError)):
        ivgan.indextts_vocoder_run(hidden[:2], conds, cond_layer, precision=capi.F32, hop=CFG.hop)


def test_session_surface(ivgan, g):
    from b200tts import session as ort
    ort.register_checkpoint("indextts_f", synth.ivgan_state(int(g["weights_seed"])))
    sess = ort.InferenceSession("IndexTTS_F.onnx", precision="fp32")
    names = [i.name for i in sess.get_inputs()]
    assert names == [f"save_bigvgan_conds_{i}" for i in range(6)] + ["bigvgan_cond_layer_speaker_embedding", "save_hidden_state"]
    assert [o.name for o in sess.get_outputs()] == ["generated_wav"]
    conds, cond_layer, hidden = synth.ivgan_inputs(int(g["a_seed"]), int(g["a_rows"]))
    feed = {f"save_bigvgan_conds_{i}": ort.OrtValue.ortvalue_from_numpy(c, "cpu", 0) for i, c in enumerate(conds)}
    feed["bigvgan_cond_layer_speaker_embedding"] = ort.OrtValue.ortvalue_from_numpy(cond_layer, "cpu", 0)
    feed["save_hidden_state"] = ort.OrtValue.ortvalue_from_numpy(hidden, "cpu", 0)
    wav = sess.run_with_ort_values(["generated_wav"], feed)[0].numpy()
    assert np.abs(wav.astype(np.int32) - g["a_pcm"].astype(np.int32)).max() <= 2

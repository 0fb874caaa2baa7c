"""GPU parity: CUDA kernels (through the C ABI) vs the oracle and the reference-made golden vectors."""
import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import capi, config, synth
from conftest import snr_db
from oracle import bigvgan_ref as R

pytestmark = pytest.mark.gpu
CFG = config.BIGVGAN


# ---- anti-aliased activation -------------------------------------------------------------------------
def test_aa_activation_golden(engine, golden_bigvgan):
    g = golden_bigvgan
    y = engine.aa_activation(g["act_x"], g["act_alpha"], g["act_beta"], g["act_filter"], precise=True, post=False)
    np.testing.assert_allclose(y, g["act_y_stage"], rtol=0, atol=2e-5)      # fp32, values O(10)
    y = engine.aa_activation(g["act_x"], g["act_alpha"], g["act_beta"], g["act_filter"], precise=True, post=True)
    assert y.shape == g["act_y_post"].shape
    np.testing.assert_allclose(y, g["act_y_post"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("C,L", [(768, 128), (384, 512), (192, 257), (96, 1000), (48, 77), (24, 4099), (24, 1), (5, 3)])
def test_aa_activation_stage_shapes(engine, C, L):
    rng = np.random.default_rng(C * 7 + L)
    x = (3.0 * rng.standard_normal((2, C, L))).astype(np.float32)
    al = (0.4 * rng.standard_normal(C)).astype(np.float32)
    be = (0.4 * rng.standard_normal(C)).astype(np.float32)
    f = R.aa_filter()
    want = R.activation1d(torch.from_numpy(x), torch.from_numpy(al), torch.from_numpy(be), f).numpy()
    got = engine.aa_activation(x, al, be, f.numpy(), precise=True, post=False)
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-5)
    fast = engine.aa_activation(x, al, be, f.numpy(), precise=False, post=False)    # __sinf variant
    np.testing.assert_allclose(fast, want, rtol=0, atol=2e-3)
    want_p = R.activation1d(torch.from_numpy(x), torch.from_numpy(al), torch.from_numpy(be), f, 15, 15, 15).numpy()
    got_p = engine.aa_activation(x, al, be, f.numpy(), precise=True, post=True)
    np.testing.assert_allclose(got_p, want_p, rtol=0, atol=3e-5)


# ---- shifted-row GEMM as Conv1d / ConvTranspose1d --------------------------------------------------------
CONV_CASES = [  # Cin, Cout, k, dil, groups, L, B
    (100, 1536, 7, 1, 1, 40, 1),      # conv_pre
    (768, 768, 3, 1, 1, 128, 2),
    (384, 384, 7, 3, 1, 200, 1),
    (192, 192, 11, 5, 1, 300, 2),
    (96, 96, 11, 1, 1, 130, 1),
    (48, 48, 7, 5, 1, 513, 1),
    (24, 24, 3, 3, 1, 1000, 2),
    (24, 24, 11, 5, 1, 7, 1),         # sequence shorter than the receptive field
    (1024, 1024, 31, 1, 16, 150, 2),  # DiT conv position embedding (grouped)
    (512, 1024, 1, 1, 1, 333, 1),     # a Linear
]


def _conv_case(Cin, Cout, k, dil, groups, L, B, seed=0):
    rng = np.random.default_rng(seed + Cin + k)
    x = rng.standard_normal((B, Cin, L)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin // groups, k)) / np.sqrt(Cin // groups * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    want = torch.nn.functional.conv1d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), dilation=dil,
                                      padding=(k * dil - dil) // 2, groups=groups).numpy()
    return x, w, b, want


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_f32(engine, case):
    Cin, Cout, k, dil, groups, L, B = case
    x, w, b, want = _conv_case(*case)
    got = engine.conv1d(x, w, b, dilation=dil, groups=groups, precision=capi.F32)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * np.sqrt(Cin // groups * k))


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_tcgen05_bf16(engine, case):
    Cin, Cout, k, dil, groups, L, B = case
    x, w, b, want = _conv_case(*case)
    got = engine.conv1d(x, w, b, dilation=dil, groups=groups, precision=capi.BF16)
    # operands rounded to bf16 (2^-9 relative), fp32 accumulate: compare against the same rounding on CPU
    xb = torch.from_numpy(x).bfloat16().float()
    wb = torch.from_numpy(w).bfloat16().float()
    want_b = torch.nn.functional.conv1d(xb, wb, torch.from_numpy(b), dilation=dil, padding=(k * dil - dil) // 2,
                                        groups=groups).numpy()
    np.testing.assert_allclose(got, want_b, rtol=0, atol=1e-4 * np.sqrt(Cin // groups * k))
    assert snr_db(want, got) > 40.0


@pytest.mark.parametrize("Cin,Cout,u,L,B", [(1536, 768, 4, 33, 1), (384, 192, 2, 200, 2), (48, 24, 2, 1000, 1), (96, 48, 2, 1, 1)])
@pytest.mark.parametrize("prec", [capi.F32, capi.BF16])
def test_conv_transpose1d(engine, Cin, Cout, u, L, B, prec):
    rng = np.random.default_rng(Cin + u)
    x = rng.standard_normal((B, Cin, L)).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, 2 * u)) / np.sqrt(Cin * 2)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    xr, wr = (x, w) if prec == capi.F32 else (torch.from_numpy(x).bfloat16().float().numpy(), torch.from_numpy(w).bfloat16().float().numpy())
    want = torch.nn.functional.conv_transpose1d(torch.from_numpy(xr), torch.from_numpy(wr), torch.from_numpy(b), stride=u,
                                                padding=u // 2).numpy()
    got = engine.conv_transpose1d(x, w, b, stride=u, precision=prec)
    assert got.shape == want.shape == (B, Cout, L * u)
    np.testing.assert_allclose(got, want, rtol=0, atol=(2e-5 if prec == capi.F32 else 1e-4) * np.sqrt(Cin * 2))


# ---- the whole graph ---------------------------------------------------------------------------------------
def test_bigvgan_f32_vs_reference_golden(bigvgan_engine, golden_bigvgan):
    """fp32 engine vs the reference's own int16 output: <= 2 LSB (fp32 re-association + truncating cast)."""
    g = golden_bigvgan
    for tag in ("a", "b"):
        pcm = bigvgan_engine.bigvgan_run(g[f"{tag}_mel"], precision=capi.F32)
        assert pcm.dtype == np.int16 and pcm.shape == g[f"{tag}_pcm"].shape
        d = np.abs(pcm.astype(np.int32) - g[f"{tag}_pcm"].astype(np.int32))
        assert d.max() <= 2, d.max()
        assert (d == 0).mean() > 0.97


def test_bigvgan_f32_vs_oracle_float(bigvgan_engine):
    sd = synth.bigvgan_state(1234)
    mel = synth.bigvgan_mel(21, 2, 40)
    pcm, wave = bigvgan_engine.bigvgan_run(mel, precision=capi.F32, return_wave=True)
    opcm, owave = R.bigvgan_pcm(mel, sd, CFG, return_float=True)
    assert np.abs(wave - owave.numpy()).max() < 0.5            # < half an LSB on a +-32767 scale
    assert np.abs(pcm.astype(np.int32) - opcm.numpy().astype(np.int32)).max() <= 1


def test_bigvgan_bf16_vs_oracle(bigvgan_engine):
    """tcgen05 path: bf16 conv operands, fp32 accumulate / residual stream. Tolerance: PCM SNR >= 30 dB."""
    sd = synth.bigvgan_state(1234)
    mel = synth.bigvgan_mel(22, 2, 48)
    pcm, wave = bigvgan_engine.bigvgan_run(mel, precision=capi.BF16, return_wave=True)
    _, owave = R.bigvgan_pcm(mel, sd, CFG, return_float=True)
    s = snr_db(owave.numpy(), wave)
    assert s > 30.0, s
    # edge samples (the +30 quirk region) carry signal, not garbage
    assert np.abs(wave[..., :15]).max() <= np.abs(owave.numpy()[..., :15]).max() * 1.5 + 50


def test_bigvgan_fp16_vs_oracle(bigvgan_engine):
    """fp16 conv operands / activations (the type BASELINE.json configs[1] names): PCM SNR >= 48 dB at this small size."""
    sd = synth.bigvgan_state(1234)
    mel = synth.bigvgan_mel(22, 2, 48)
    pcm, wave = bigvgan_engine.bigvgan_run(mel, precision=capi.F16, return_wave=True)
    _, owave = R.bigvgan_pcm(mel, sd, CFG, return_float=True)
    s = snr_db(owave.numpy(), wave)
    assert s > 48.0, s
    assert np.isfinite(wave).all()


@pytest.mark.parametrize("case", [(768, 768, 7, 3, 1, 2, 300), (24, 24, 11, 5, 1, 1, 4099), (1024, 1024, 31, 1, 16, 2, 130)])
def test_conv1d_tcgen05_fp16(engine, case):
    Cin, Cout, k, dil, groups, B, L = case
    rng = np.random.default_rng(Cin + k)
    x = rng.standard_normal((B, Cin, L)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin // groups, k)) / np.sqrt(Cin // groups * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    got = engine.conv1d(x, w, b, dilation=dil, groups=groups, precision=capi.F16)
    xh, wh = torch.from_numpy(x).half().float(), torch.from_numpy(w).half().float()
    want = torch.nn.functional.conv1d(xh, wh, torch.from_numpy(b), dilation=dil, padding=(k * dil - dil) // 2, groups=groups).numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4 * np.sqrt(Cin // groups * k))


def test_bigvgan_batch_items_independent(bigvgan_engine):
    mel = synth.bigvgan_mel(23, 3, 16)
    all_ = bigvgan_engine.bigvgan_run(mel, precision=capi.F32)
    one = bigvgan_engine.bigvgan_run(mel[1:2], precision=capi.F32)
    np.testing.assert_array_equal(all_[1:2], one)


def test_bigvgan_session_surface(bigvgan_engine):
    """The reference's call sequence (BigVGAN/Export_BigVGAN.py:153-174) against the drop-in session."""
    from b200tts import session as onnxruntime
    onnxruntime._engines[0] = bigvgan_engine
    onnxruntime.register_checkpoint("bigvgan", synth.bigvgan_state(1234))
    sess = onnxruntime.InferenceSession("/tmp/BigVGAN.onnx", sess_options=onnxruntime.SessionOptions(), providers=[],
                                        provider_options=None, precision="fp32")
    assert "float16" not in sess._inputs_meta[0].type
    in_name, out_name = sess.get_inputs()[0].name, sess.get_outputs()[0].name
    assert (in_name, out_name) == ("mel_features", "generated_wav")
    dummy = onnxruntime.OrtValue.ortvalue_from_numpy(np.ones((1, sess._inputs_meta[0].shape[1], 8), dtype=np.float32), "cpu", 0)
    out = sess.run_with_ort_values([out_name], {in_name: dummy})
    wav = out[0].numpy()
    assert wav.dtype == np.int16 and wav.shape == (1, 1, 8 * 256 + 30)
    want = R.bigvgan_pcm(np.ones((1, 100, 8), dtype=np.float32), synth.bigvgan_state(1234), CFG).numpy()
    assert np.abs(wav.astype(np.int32) - want).max() <= 2
    with pytest.raises(ValueError):
        sess.run([out_name], {in_name: np.ones((1, 100, 0), dtype=np.float32)})


def test_bigvgan_fp16_session_reports_float16_and_takes_it(bigvgan_engine):
    """The fp16 graph variant (BASELINE.json configs[1], q6): the session reports a float16 input exactly as the reference's fp16
    export does, so the reference's own dtype switch (Export_BigVGAN.py:155-165) feeds float16 mels; the PCM stays within the fp16
    engine's bar of the fp32 reference."""
    from b200tts import session as onnxruntime
    onnxruntime._engines[0] = bigvgan_engine
    onnxruntime.register_checkpoint("bigvgan", synth.bigvgan_state(1234))
    sess = onnxruntime.InferenceSession("/tmp/BigVGAN.onnx", providers=[], precision="fp16")
    model_dtype = sess._inputs_meta[0].type
    assert "float16" in model_dtype
    mel = synth.bigvgan_mel(11, 1, 24)
    feed = onnxruntime.OrtValue.ortvalue_from_numpy(mel.astype(np.float16), "cpu", 0)
    wav = sess.run_with_ort_values(["generated_wav"], {"mel_features": feed})[0].numpy()
    want = R.bigvgan_pcm(mel.astype(np.float16).astype(np.float32), synth.bigvgan_state(1234), CFG).numpy()
    assert wav.dtype == np.int16 and wav.shape == want.shape
    assert snr_db(want, wav) > 45.0

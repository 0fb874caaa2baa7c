"""GPU parity for the F5-TTS path: CUDA engine (through the C ABI / the drop-in sessions) vs the vectors made by the
reference's own F5Preprocess / F5Transformer / F5Decode (tests/golden/f5_ref.npz) and vs the oracle at other sizes.

Stated tolerances (confirmed empirically on B200):
  fp32 engine : mel max-abs <= 1e-3 after all 31 Euler steps; PCM within a few LSB of the reference's int16
  bf16 engine : mel cosine >= 0.999 after 31 steps; PCM SNR >= 25 dB vs the fp32 reference (N = 130: short, noisy utterance)
  fp16 engine : mel cosine >= 0.99999 after 31 steps; PCM SNR >= 45 dB (the benchmarked operand type; the BASELINE-size
                bars are in test_gpu_fullsize.py)
"""
import os

import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import capi, config, synth, weights
from conftest import GOLDEN, snr_db
from oracle import f5_ref as R

pytestmark = pytest.mark.gpu
CFG = config.F5


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "f5_ref.npz")))


@pytest.fixture(scope="module")
def f5(engine, g):
    dsd = synth.f5_dit_state(int(g["dit_seed"]))
    engine.load_state("dit", weights.dit_engine_tensors(dsd, CFG))
    engine.load_state("vocos", weights.vocos_engine_tensors(synth.vocos_state(int(g["vocos_seed"])), CFG))
    engine.load_state("f5", weights.f5_export_constants(dsd, CFG))
    engine.f5_build()
    return engine


def _inputs(g):
    return synth.f5_inputs(int(g["input_seed"]), int(g["audio_len"]), int(g["n_text"]))


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


# ---- attention kernel alone -----------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 130, 1126])
def test_attention_tcgen05_fp16(engine, N):
    """fp16 operands: q, k, v and the probabilities carry 11 significant bits -> tighter than the bf16 bar below."""
    rng = np.random.default_rng(N + 7)
    H = 16
    q = (0.35 * rng.standard_normal((2, H, N, 64))).astype(np.float32)
    k = (0.35 * rng.standard_normal((2, H, N, 64))).astype(np.float32)
    v = rng.standard_normal((2, H, N, 64)).astype(np.float32)
    got = engine.attention(q, k, v, precision=capi.F16)
    qh, kh, vh = (torch.from_numpy(t).half().float() for t in (q, k, v))
    want = torch.softmax(qh @ kh.transpose(-1, -2), dim=-1) @ vh                    # (2, H, N, 64), fp32 math on fp16-rounded operands
    want = want.permute(0, 2, 1, 3).reshape(2, N, H * 64).numpy()
    assert np.abs(got - want).max() <= 4e-3 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("N", [1, 64, 128, 130, 257, 1126])
def test_attention_tcgen05(engine, N):
    rng = np.random.default_rng(N)
    H = 16
    q, k, v = (rng.standard_normal((2, H, N, 64)).astype(np.float32) for _ in range(3))
    q *= 0.35
    k *= 0.35                                     # logits ~ N(0, 1): the regime the pre-scaled weights produce
    got = engine.attention(q, k, v)
    qb, kb, vb = (torch.from_numpy(t).bfloat16().float() for t in (q, k, v))
    want = torch.matmul(torch.softmax(torch.matmul(qb, kb.transpose(-1, -2)), dim=-1), vb).transpose(1, 2).reshape(2, N, H * 64)
    # P and the output are rounded to bf16 (2^-9 relative); accumulation is fp32
    np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=2.5e-2)
    assert cosine(got, want.numpy()) > 0.9995


# ---- graph A ------------------------------------------------------------------------------------------------
def test_preprocess_vs_reference(f5, g):
    audio, text_ids, maxd, _ = _inputs(g)
    cond, cond_drop, ref_len = f5.f5_preprocess(audio, text_ids, int(maxd[0]))
    assert ref_len == int(g["ref_signal_len"])
    assert cond.shape == g["cat_mel_text"].shape == (1, 130, 612)
    # log-mel half: STFT as an fp32 GEMM re-associates sums; log() amplifies that only where |X| is tiny
    np.testing.assert_allclose(cond[..., :100], g["cat_mel_text"][..., :100], rtol=0, atol=5e-3)
    assert np.abs(cond[..., :100] - g["cat_mel_text"][..., :100]).mean() < 1e-4
    np.testing.assert_allclose(cond[..., 100:], g["cat_mel_text"][..., 100:], rtol=0, atol=2e-4)
    np.testing.assert_allclose(cond_drop[..., 100:], g["cat_mel_text_drop"][..., 100:], rtol=0, atol=2e-4)
    assert np.abs(cond_drop[..., :100]).max() == 0.0
    nt = int(g["n_text"])
    assert np.abs(cond[0, nt:, 100:]).max() == 0.0 and np.abs(cond_drop[0, nt:, 100:]).max() == 0.0   # masked filler rows


# ---- graph B ------------------------------------------------------------------------------------------------
def test_transformer_f32_steps_vs_reference(f5, g):
    _, _, _, noise = _inputs(g)
    x, ts = noise, 0
    for want_ts in (1, 2):
        x, ts = f5.f5_transformer(x, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], ts,
                                  n_steps=1, precision=capi.F32)
        assert ts == want_ts
        np.testing.assert_allclose(x, g[f"noise_after_{ts}"], rtol=0, atol=2e-5)
    x, ts = f5.f5_transformer(x, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], ts,
                              n_steps=6, precision=capi.F32)
    np.testing.assert_allclose(x, g["noise_after_8"], rtol=0, atol=1e-4)
    x, ts = f5.f5_transformer(x, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], ts,
                              n_steps=23, precision=capi.F32)
    assert ts == 31
    assert np.abs(x - g["noise_after_31"]).max() <= 1e-3          # the stated fp32 tolerance after 31 Euler steps
    with pytest.raises(RuntimeError, match="time_step out of range"):
        f5.f5_transformer(x, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], ts, 1, capi.F32)


def test_transformer_bf16_vs_reference(f5, g):
    _, _, _, noise = _inputs(g)
    x1, _ = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0,
                              n_steps=1, precision=capi.BF16)
    # one step: the update is delta_t[0] ~ 1.3e-3 times the prediction, so compare the prediction itself
    dt0 = float(g["delta_t"][0])
    pred_ref = (g["noise_after_1"] - noise) / dt0
    pred_got = (x1 - noise) / dt0
    assert cosine(pred_got, pred_ref) > 0.995
    x, ts = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0,
                              n_steps=31, precision=capi.BF16)
    assert ts == 31
    assert cosine(x, g["noise_after_31"]) >= 0.999                # the stated bf16 tolerance
    # fused n_steps == loop of single steps, bit for bit (same kernels, same order)
    y, t2 = noise, 0
    for _ in range(3):
        y, t2 = f5.f5_transformer(y, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], t2, 1, capi.BF16)
    z, _ = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0, 3, capi.BF16)
    np.testing.assert_array_equal(y, z)


def test_transformer_fp16_and_fused_chain_vs_reference(f5, g):
    """fp16 operands (kind::f16 with the fp16 format code) and the fused row-block chain (dit_chain.cu) against the reference's
    vectors; chain on / off differ only by the LayerNorm variance formula (E[x^2] - mean^2 in the fused epilogue)."""
    _, _, _, noise = _inputs(g)
    dt0 = float(g["delta_t"][0])
    pred_ref = (g["noise_after_1"] - noise) / dt0
    out = {}
    try:
        for prec in (capi.F16, capi.BF16):
            for chain in (0, 1):
                f5.set_option("dit_chain", chain)
                x1, _ = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0,
                                          n_steps=1, precision=prec)
                assert cosine((x1 - noise) / dt0, pred_ref) > (0.9999 if prec == capi.F16 else 0.995)
                x, ts = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0,
                                          n_steps=31, precision=prec)
                assert ts == 31 and np.isfinite(x).all()
                assert cosine(x, g["noise_after_31"]) >= (0.99999 if prec == capi.F16 else 0.999)
                out[(prec, chain)] = x
            # the two code paths agree far more closely with each other than either does with the fp32 reference
            d_paths = np.abs(out[(prec, 0)] - out[(prec, 1)]).max()
            d_ref = np.abs(out[(prec, 0)] - g["noise_after_31"]).max()
            assert d_paths <= max(2.0 * d_ref, 1e-3), (d_paths, d_ref)
        # fused chain: n fused steps == n single-step calls, bit for bit
        f5.set_option("dit_chain", 1)
        y, t2 = noise, 0
        for _ in range(3):
            y, t2 = f5.f5_transformer(y, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], t2, 1, capi.F16)
        z, _ = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0, 3, capi.F16)
        np.testing.assert_array_equal(y, z)
    finally:
        f5.set_option("dit_chain", 1)


def test_fused_chain_one_step(f5, g):
    """One Euler step at N = 130 through the fused chain (two row blocks, team of 8) == the seven-launch path to fp16 noise.
    Small on purpose: this is the case tools/sanitize.sh runs under compute-sanitizer racecheck / synccheck."""
    _, _, _, noise = _inputs(g)
    out = {}
    try:
        for chain in (0, 1):
            f5.set_option("dit_chain", chain)
            out[chain], _ = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0,
                                              n_steps=1, precision=capi.F16)
    finally:
        f5.set_option("dit_chain", 1)
    dt0 = float(g["delta_t"][0])
    assert cosine((out[1] - noise) / dt0, (out[0] - noise) / dt0) > 0.99999
    assert cosine((out[1] - noise) / dt0, (g["noise_after_1"] - noise) / dt0) > 0.9999


def test_fused_chain_one_step_fp8(f5, g):
    """The same step with the e4m3 option (ff1 and q|k|v on kind::f8f6f4): the predicted velocity stays within e4m3's error of the
    reference's (cosine > 0.998; 0.99999 without the option). Also a compute-sanitizer case (tools/sanitize.sh)."""
    _, _, _, noise = _inputs(g)
    dt0 = float(g["delta_t"][0])
    try:
        for level in (1, 2):
            f5.set_option("dit_fp8", level)
            out, _ = f5.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0, n_steps=1,
                                       precision=capi.F16)
            c = cosine((out - noise) / dt0, (g["noise_after_1"] - noise) / dt0)
            assert np.isfinite(out).all() and 0.997 < c < 0.99999, (level, c)
    finally:
        f5.set_option("dit_fp8", 0)


def test_synthesize_fp16_vs_reference(f5, g):
    audio, text_ids, maxd, noise = _inputs(g)
    pcm, mel = f5.f5_synthesize(audio, text_ids, int(maxd[0]), noise, precision=capi.F16, return_mel=True)
    assert cosine(mel, g["noise_after_31"]) >= 0.99999
    assert snr_db(g["pcm"], pcm) > 45.0


# ---- graph C ------------------------------------------------------------------------------------------------
def test_decode_vs_reference(f5, g):
    pcm, wave = f5.f5_decode(g["decode_in"], 12, return_wave=True)
    assert pcm.dtype == np.int16 and pcm.shape == g["decode_pcm"].shape
    d = np.abs(pcm.astype(np.int32) - g["decode_pcm"])
    assert d.max() <= 3 and (d <= 1).mean() > 0.99
    pcm = f5.f5_decode(g["noise_after_31"], int(g["ref_signal_len"]))
    d = np.abs(pcm.astype(np.int32) - g["pcm"])
    assert d.max() <= 3 and (d <= 1).mean() > 0.99
    with pytest.raises(RuntimeError, match="ref_signal_len"):
        f5.f5_decode(g["decode_in"], 40)
    assert f5.f5_decode(g["decode_in"], 39).shape == (1, 1, 0)    # a single frame decodes to 256*(1-1) = 0 samples


# ---- A + 31 x B + C ------------------------------------------------------------------------------------------
def test_synthesize_vs_reference(f5, g):
    audio, text_ids, maxd, noise = _inputs(g)
    pcm32, mel32 = f5.f5_synthesize(audio, text_ids, int(maxd[0]), noise, precision=capi.F32, return_mel=True)
    assert pcm32.shape == g["pcm"].shape
    assert np.abs(mel32 - g["noise_after_31"]).max() <= 1e-3
    assert snr_db(g["pcm"], pcm32) > 55.0
    pcm16, mel16 = f5.f5_synthesize(audio, text_ids, int(maxd[0]), noise, precision=capi.BF16, return_mel=True)
    assert cosine(mel16, g["noise_after_31"]) >= 0.999
    assert snr_db(g["pcm"], pcm16) > 25.0


def test_synthesize_other_sizes_vs_oracle(f5, g):
    """Ragged sizes (N not a multiple of anything, ref != gen length) against the oracle run here on the CPU."""
    dsd, vsd = synth.f5_dit_state(int(g["dit_seed"])), synth.vocos_state(int(g["vocos_seed"]))
    audio, text_ids, _, _ = synth.f5_inputs(5, audio_len=9000, n_text=13)
    N = 36 + 41
    noise = np.random.default_rng(9).standard_normal((1, N, 100), dtype=np.float32)
    want_pcm, want_mel, ref_len = R.f5_synthesize(audio, text_ids, [N], noise, dsd, vsd, CFG, steps=4, return_mel=True)
    pcm, mel = f5.f5_synthesize(audio, text_ids, N, noise, precision=capi.F32, n_steps=4, return_mel=True)
    assert ref_len == 9000 // 256 + 1 and pcm.shape == tuple(want_pcm.shape)
    np.testing.assert_allclose(mel, want_mel.numpy(), rtol=0, atol=2e-4)
    assert snr_db(want_pcm.numpy(), pcm) > 55.0


def test_synthesize_batch_matches_single(f5, g):
    """Length-bucketed batching (config 4): U utterances through ONE batched DiT loop give what U single calls give."""
    import torch
    U, L, n_text = 3, 9000, 20
    ins = [synth.f5_inputs(50 + i, L, n_text) for i in range(U)]
    N = int(ins[0][2][0])
    ns = 256 * (N - (L // 256 + 1) - 1)
    audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
    ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
    noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
    pcm_b = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
    mel_b = torch.zeros((U, N, CFG.n_mels), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    f5.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm_b.data_ptr(),
                                  precision=capi.BF16, n_steps=4, mel_ptr=mel_b.data_ptr())
    f5.synchronize()
    for u in range(U):
        pcm_1 = torch.zeros((ns,), dtype=torch.int16, device="cuda")
        mel_1 = torch.zeros((N, CFG.n_mels), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        f5.f5_synthesize_device(audio[u].data_ptr(), L, ids[u].data_ptr(), n_text, N, noise[u].data_ptr(), pcm_1.data_ptr(),
                                precision=capi.BF16, n_steps=4, mel_ptr=mel_1.data_ptr())
        f5.synchronize()
        # every kernel is row-wise or per-sequence and accumulates over K in the same order whatever the tile shape
        np.testing.assert_allclose(mel_b[u].cpu().numpy(), mel_1.cpu().numpy(), rtol=0, atol=1e-5)
        assert np.abs(pcm_b[u].cpu().numpy().astype(np.int32) - pcm_1.cpu().numpy().astype(np.int32)).max() <= 1
    assert float(mel_b.abs().max()) > 0.1 and not bool(torch.equal(mel_b[0], mel_b[1]))


def test_f5_session_surface(f5, g):
    """The reference's own loop (F5-TTS-ONNX-Inference.py:247-311) against the drop-in sessions."""
    from b200tts import session as onnxruntime
    onnxruntime._engines[0] = f5
    onnxruntime._f5_ready[id(f5)] = True
    onnxruntime.set_seed(9527)
    opts = onnxruntime.SessionOptions()
    A = onnxruntime.InferenceSession("F5_Preprocess.onnx", sess_options=opts, providers=["CPUExecutionProvider"])
    B = onnxruntime.InferenceSession("F5_Transformer.onnx", sess_options=opts, providers=[], precision="fp32")
    C = onnxruntime.InferenceSession("F5_Decode.onnx", sess_options=opts, providers=["CPUExecutionProvider"])
    assert [i.name for i in A.get_inputs()] == ["audio", "text_ids", "max_duration"]
    assert [o.name for o in A.get_outputs()] == ["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k",
                                                 "cat_mel_text", "cat_mel_text_drop", "ref_signal_len"]
    assert [i.name for i in B.get_inputs()][-1] == "time_step" and [o.name for o in B.get_outputs()] == ["denoised", "time_step"]
    audio, text_ids, maxd, noise_in = _inputs(g)
    outs = A.run([o.name for o in A.get_outputs()], {"audio": audio, "text_ids": text_ids, "max_duration": maxd})
    noise, rcq, rsq, rck, rsk, cmt, cmtd, ref_len = outs
    assert noise.shape == (1, 130, 100) and rcq.shape == (2, 16, 130, 64) and rck.shape == (2, 16, 64, 130)
    np.testing.assert_array_equal(rcq[1, 7], g["rope_cos_row"])
    noise = noise_in                                   # parity: inject the Euler start (ORT's RNG is irreproducible)
    time_step = np.array([0], dtype=np.int32)
    in_B = [i.name for i in B.get_inputs()]
    for _ in range(0, 32 - 1, 1):
        noise, time_step = B.run(["denoised", "time_step"], dict(zip(in_B, [noise, rcq, rsq, rck, rsk, cmt, cmtd, time_step])))
    assert int(time_step[0]) == 31
    assert np.abs(noise - g["noise_after_31"]).max() <= 1e-3
    fused, ts = B.run_all_steps(dict(zip(in_B, [noise_in, rcq, rsq, rck, rsk, cmt, cmtd, np.array([0], dtype=np.int32)])))
    np.testing.assert_array_equal(fused, noise)        # per-step calls stay bit-compatible with the fused path
    wav = C.run(["output_audio"], {"denoised": noise, "ref_signal_len": ref_len})[0]
    assert wav.dtype == np.int16 and wav.shape == g["pcm"].shape
    assert snr_db(g["pcm"], wav) > 55.0


def test_f5_iobinding_loop(f5, g):
    """The reference's bound loop (F5-TTS-ONNX-Inference.py:257-288): outputs aliased onto the inputs, 31 bound runs."""
    from b200tts import session as ort
    ort._engines[0] = f5
    ort._f5_ready[id(f5)] = True
    B = ort.InferenceSession("F5_Transformer.onnx", precision="fp32")
    in_names = [i.name for i in B.get_inputs()]
    out_names = [o.name for o in B.get_outputs()]
    audio, ids, maxd, noise = _inputs(g)
    N = int(maxd[0])
    rq = np.broadcast_to(g["rope_cos_row"], (2, 16, N, 64)).copy()
    sq = np.broadcast_to(g["rope_sin_row"], (2, 16, N, 64)).copy()
    feed = [noise.copy(), rq, sq, rq.transpose(0, 1, 3, 2).copy(), sq.transpose(0, 1, 3, 2).copy(),
            g["cat_mel_text"], g["cat_mel_text_drop"], np.array([0], dtype=np.int32)]
    inputs = [ort.OrtValue.ortvalue_from_numpy(a, "cuda", 0) for a in feed]
    outputs = [inputs[0], inputs[-1]]
    io = B.io_binding()
    for name, v in zip(in_names, inputs):
        io.bind_ortvalue_input(name=name, ortvalue=v)
    for name, v in zip(out_names, outputs):
        io.bind_ortvalue_output(name=name, ortvalue=v)
    for _ in range(31):
        B.run_with_iobinding(io)
    got = ort.OrtValue.numpy(io.get_outputs()[0])
    assert int(inputs[-1].numpy()[0]) == 31
    assert np.abs(got - g["noise_after_31"]).max() <= 1e-3


# ---- ragged batches (config 4: utterances of different lengths share one DiT loop) -----------------------------------------
def test_ragged_batch_equals_single_utterances(f5, g):
    """Three utterances of different (audio length, text length, max_duration) through b200tts_f5_bigvgan_pipeline_ragged ==
    each of them alone (the reference runs one utterance per script run, F5-TTS-ONNX-Inference.py:227-311). Rows of a sequence
    sit at other offsets of the row blocks / attention tiles in the batch; a row's result must not depend on that. One of them
    is also held against the oracle on the CPU."""
    from b200tts import weights as W
    f5.load_state("bigvgan", W.bigvgan_engine_tensors(synth.bigvgan_state(1234)))
    f5.bigvgan_build()
    specs = [(9000, 13, 36 + 41), (16384, 20, 130), (12345, 17, 49 + 70)]          # (L, n_text, N): ref frames L//256+1 = 36, 65, 49
    audios, texts, Ns, noises = [], [], [], []
    for i, (L, nt, N) in enumerate(specs):
        a, t, _, _ = synth.f5_inputs(50 + i, audio_len=L, n_text=nt)
        audios.append(a.reshape(-1)); texts.append(t.reshape(-1)); Ns.append(N)
        noises.append(np.random.default_rng(90 + i).standard_normal((N, 100), dtype=np.float32))
    for prec, steps in ((capi.F16, 31), (capi.BF16, 4)):
        wavs, vocs, mels = f5.f5_bigvgan_pipeline_ragged(audios, texts, Ns, noises, precision=prec, n_steps=steps, with_vocos=True, return_mel=True)
        for u, (L, nt, N) in enumerate(specs):
            G = N - (L // 256 + 1)
            assert wavs[u].shape == (256 * G + 30,) and vocs[u].shape == (256 * (G - 1),) and mels[u].shape == (N, 100)
            w1, v1, m1 = f5.f5_bigvgan_pipeline(audios[u][None], texts[u][None], N, noises[u][None], precision=prec, n_steps=steps,
                                               with_vocos=True, return_mel=True)
            np.testing.assert_allclose(mels[u], m1[0], rtol=0, atol=2e-5)
            assert snr_db(w1[0], wavs[u]) >= 60.0 and snr_db(v1[0], vocs[u]) >= 60.0
    # fp16, all 31 steps, utterance 0 against the oracle
    dsd, vsd = synth.f5_dit_state(int(g["dit_seed"])), synth.vocos_state(int(g["vocos_seed"]))
    L, nt, N = specs[0]
    wavs, vocs, mels = f5.f5_bigvgan_pipeline_ragged(audios, texts, Ns, noises, precision=capi.F16, with_vocos=True, return_mel=True)
    want_pcm, want_mel, _ = R.f5_synthesize(audios[0].reshape(1, 1, -1), texts[0].reshape(1, -1), [N], noises[0][None], dsd, vsd, CFG,
                                            steps=31, return_mel=True)
    assert cosine(mels[0], want_mel.numpy()) >= 0.99999
    assert snr_db(want_pcm.numpy().reshape(-1), vocs[0]) > 45.0


def test_ragged_embed_option_changes_nothing(f5, g):
    """b200tts_set_option("ragged_embed", 1): the input embedding of a ragged batch as ONE launch over all DiT rows (x gathered
    into the row order) instead of two launches per utterance -- the same numbers (a row's dot products do not depend on the tile
    it lands in), and still equal to single utterances."""
    from b200tts import weights as W
    f5.load_state("bigvgan", W.bigvgan_engine_tensors(synth.bigvgan_state(1234)))
    f5.bigvgan_build()
    specs = [(9000, 13, 36 + 41), (16384, 20, 130), (12345, 17, 49 + 70), (7000, 9, 61)]
    audios, texts, Ns, noises = [], [], [], []
    for i, (L, nt, N) in enumerate(specs):
        a, t, _, _ = synth.f5_inputs(150 + i, audio_len=L, n_text=nt)
        audios.append(a.reshape(-1)); texts.append(t.reshape(-1)); Ns.append(N)
        noises.append(np.random.default_rng(190 + i).standard_normal((N, 100), dtype=np.float32))
    runs = {}
    try:
        for flag in (0, 1):
            f5.set_option("ragged_embed", flag)
            runs[flag] = f5.f5_bigvgan_pipeline_ragged(audios, texts, Ns, noises, precision=capi.F16, n_steps=6, with_vocos=True, return_mel=True)
        f5.set_option("ragged_embed", 1)
        single = f5.f5_bigvgan_pipeline(audios[3][None], texts[3][None], Ns[3], noises[3][None], precision=capi.F16, n_steps=6,
                                        with_vocos=True, return_mel=True)
    finally:
        f5.set_option("ragged_embed", 0 if os.environ.get("B200TTS_RAGGED_EMBED", "1") == "0" else 1)        # the default is on
    for u in range(len(specs)):
        np.testing.assert_allclose(runs[1][2][u], runs[0][2][u], rtol=0, atol=2e-5)
        assert snr_db(runs[0][0][u], runs[1][0][u]) >= 60.0 and snr_db(runs[0][1][u], runs[1][1][u]) >= 60.0
    np.testing.assert_allclose(runs[1][2][3], single[2][0], rtol=0, atol=2e-5)


def test_frontend_wav_and_text_in_wav_out(f5, g, tmp_path):
    """The reference script's call surface (F5-TTS-ONNX-Inference.py:13-39,223-315): a wav file + reference text + text to speak in,
    a wav file out. The front end's ids / duration / noise drive the same engine call a direct caller would make."""
    from b200tts import frontend as fe
    import string
    symbols = [" "] + list(string.ascii_lowercase) + list(",.'!?")
    vocab_path = tmp_path / "vocab.txt"
    vocab_path.write_text("".join(s + "\n" for s in symbols), encoding="utf-8")
    audio, _, _, _ = synth.f5_inputs(7, audio_len=24000, n_text=4)
    ref_wav = str(tmp_path / "ref.wav")
    fe.save_wav(ref_wav, audio.reshape(-1))
    tts = fe.F5Synthesizer(f5, str(vocab_path), precision=capi.F16, nfe_steps=4)
    ref_text, gen_text = "hello there, world.", "this is a test."
    out_wav = str(tmp_path / "generated.wav")
    pcm = tts.synthesize(ref_wav, ref_text, gen_text, out_path=out_wav)
    a, ids, maxd, noise = tts.prepare(ref_wav, ref_text, gen_text)
    frames = 24000 // 256 + 1
    assert maxd == frames + int(frames / len(ref_text) * len(gen_text))
    assert ids.shape == (1, len(ref_text + gen_text) + 1)          # 'world.' + 'this' glued: the symbol pass inserts one space
    np.testing.assert_array_equal(a.reshape(-1), audio.reshape(-1))
    want = f5.f5_synthesize(a, ids, maxd, noise, precision=capi.F16, n_steps=4).reshape(-1)
    np.testing.assert_array_equal(pcm, want)
    assert pcm.size == 256 * (maxd - frames - 1)
    np.testing.assert_array_equal(fe.load_wav_mono_int16(out_wav), pcm)

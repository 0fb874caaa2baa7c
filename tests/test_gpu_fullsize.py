"""GPU, BASELINE.json's FULL sizes.

Against the reference itself: tests/golden/fullsize_ref.npz holds what the reference's own modules produce at configs[0..2]
sizes (oracle/make_golden_fullsize.py: BigVGAN (1,100,512); F5 N = 1126, all 31 steps, graph C). At these sizes the engine
picks other tiles than the small tests do (multi-wave schedules, the CTA-pair GEMM, the fused DiT chain with 9 teams).

Stated tolerances at these sizes (measured on B200, profiles/r02/parity_fullsize.md):
  fp32 engine          BigVGAN PCM <= 2 LSB;  F5 mel max-abs <= 1e-3 after 31 steps, PCM <= 3 LSB
  fp16 engine (bench)  BigVGAN PCM SNR >= 50 dB (57.1 measured);  F5 mel cosine >= 0.99999, PCM SNR >= 55 dB (61.9 measured)
  bf16 engine          BigVGAN PCM SNR >= 36 dB (39.0 measured);  F5 mel cosine >= 0.9999,  PCM SNR >= 40 dB (44.0 measured)

Plus size-independent properties of the path:

  * batch decomposition: a batch equals its items run alone (the reference graphs are batch-1) -- bit-exact,
  * tiling invariance: the first frames of a long mel do not depend on what follows beyond the receptive field,
  * shift behaviour of the vocoder: interior PCM of a mel shifted by k frames is the PCM shifted by 256 k samples,
  * determinism of repeated / graph-replayed calls,
  * batched F5 utterances equal single utterances at config-3 size (M = 2*U*1126 rows pick different GEMM tiles, the
    CTA-pair kernel and multi-wave schedules, none of which may change a row's result),
  * every output is finite and uses the int16 range sensibly.
"""
import ctypes

import numpy as np
import pytest
import torch

import os

import b200tts  # noqa: F401
from b200tts import capi, config, synth, weights
from conftest import GOLDEN, snr_db

pytestmark = pytest.mark.gpu
VCFG, FCFG = config.BIGVGAN, config.F5


@pytest.fixture(scope="module")
def gf():
    return dict(np.load(os.path.join(GOLDEN, "fullsize_ref.npz")))


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def test_bigvgan_config1_vs_reference_golden(bigvgan_engine, gf):
    """configs[0]/[1] size, mel (1,100,512): fp32 engine vs the reference's int16 within 2 LSB; the 16-bit engines by PCM SNR,
    alone and as an item of a batch of 8 (configs[1]: CTA-pair convolutions, 3-wave schedules)."""
    mel = synth.bigvgan_mel(int(gf["vgan_mel_seed"]), 1, int(gf["vgan_T"]))
    want = gf["vgan_pcm"].astype(np.int32)
    got = bigvgan_engine.bigvgan_run(mel, precision=capi.F32).astype(np.int32)
    assert got.shape == want.shape == (1, 1, 131102)
    d = np.abs(got - want)
    assert d.max() <= 2 and (d <= 1).mean() > 0.999
    mel8 = synth.bigvgan_mel(100, 8, 512)
    mel8[3] = mel[0]
    for prec, floor in ((capi.F16, 50.0), (capi.BF16, 36.0)):
        one = bigvgan_engine.bigvgan_run(mel, precision=prec)
        assert snr_db(want, one) >= floor, (prec, snr_db(want, one))
        batch = bigvgan_engine.bigvgan_run(mel8, precision=prec)
        np.testing.assert_array_equal(batch[3], one[0])




def test_bigvgan_config2_batch_equals_items(bigvgan_engine):
    """configs[1]: mels (8,100,512) bf16 -- the batch run (CTA-pair convs, 3-wave schedules) equals eight batch-1 runs."""
    mel = synth.bigvgan_mel(100, 8, 512)
    full = bigvgan_engine.bigvgan_run(mel, precision=capi.BF16)
    assert full.shape == (8, 1, VCFG.out_samples(512)) and full.dtype == np.int16
    for i in (0, 3, 7):
        one = bigvgan_engine.bigvgan_run(mel[i:i + 1], precision=capi.BF16)
        np.testing.assert_array_equal(full[i:i + 1], one)
    again = bigvgan_engine.bigvgan_run(mel, precision=capi.BF16)          # second call replays the captured graph
    np.testing.assert_array_equal(full, again)
    rms = np.sqrt((full.astype(np.float64) ** 2).mean())
    assert 200 < rms < 20000 and np.abs(full).max() <= 32767


def test_bigvgan_concurrent_branches_equal_serial(bigvgan_engine):
    """The three resblocks of a stage run as concurrent branches (engine option bigvgan_branches, on by default) and add into the
    stage output in the reference's order: the PCM is bit-identical to the serial schedule, for a batch and for one utterance,
    eager and graph-replayed."""
    try:
        for B, T in ((8, 512), (1, 563), (2, 37)):
            mel = synth.bigvgan_mel(300 + T, B, T)
            outs = {}
            for br in (0, 1):
                bigvgan_engine.set_option("bigvgan_branches", br)
                first = bigvgan_engine.bigvgan_run(mel, precision=capi.F16)
                again = bigvgan_engine.bigvgan_run(mel, precision=capi.F16)
                third = bigvgan_engine.bigvgan_run(mel, precision=capi.F16)       # eager, captured, replayed
                np.testing.assert_array_equal(first, again)
                np.testing.assert_array_equal(first, third)
                outs[br] = first
            np.testing.assert_array_equal(outs[0], outs[1])
    finally:
        bigvgan_engine.set_option("bigvgan_branches", 1)


def test_bigvgan_prefix_and_shift_invariance(bigvgan_engine):
    """Receptive field of the generator is finite (about 40 mel frames either side: the k = 11, dilation 5 resblocks of the
    first stage dominate): the PCM of frames [0, 256) does not change when more frames follow 64 frames later, and a mel
    shifted by 16 frames gives the PCM shifted by 4096 samples in the interior."""
    mel = synth.bigvgan_mel(101, 1, 512)
    full = bigvgan_engine.bigvgan_run(mel, precision=capi.F32).astype(np.int32)
    head = bigvgan_engine.bigvgan_run(mel[:, :, :320], precision=capi.F32).astype(np.int32)
    n = 15 + 256 * 256                                                       # samples of the first 256 frames (+ the 15-sample lead)
    assert np.abs(full[..., :n] - head[..., :n]).max() <= 1
    k, margin = 16, 64
    shifted = bigvgan_engine.bigvgan_run(mel[:, :, k:], precision=capi.F32).astype(np.int32)
    lo, hi = 15 + 256 * margin, 15 + 256 * (512 - k - margin)                # interior of the shifted signal
    assert np.abs(shifted[..., lo:hi] - full[..., lo + 256 * k:hi + 256 * k]).max() <= 1


@pytest.fixture(scope="module")
def f5_engine(engine):
    dsd = synth.f5_dit_state(4321)
    engine.load_state("dit", weights.dit_engine_tensors(dsd, FCFG))
    engine.load_state("vocos", weights.vocos_engine_tensors(synth.vocos_state(2468), FCFG))
    engine.load_state("f5", weights.f5_export_constants(dsd, FCFG))
    engine.f5_build()
    return engine


def test_f5_config3_vs_reference_golden(f5_engine, gf):
    """configs[2]: 6 s reference, 150 text ids, N = 1126, 31 Euler steps, Vocos / ISTFT decode -- against the reference's own
    graphs A, B (x31), C. fp32 engine to the fp32 tolerance; fp16 (the benchmarked type) and bf16 engines, fused chain on and off."""
    audio, ids, maxd, noise = synth.f5_inputs(int(gf["input_seed"]), int(gf["audio_len"]), int(gf["n_text"]))
    N = int(maxd[0])
    ref_len = int(gf["f5_ref_signal_len"])
    assert N == 1126 and ref_len == 563
    pcm, mel = f5_engine.f5_synthesize(audio, ids, N, noise, precision=capi.F32, return_mel=True)
    assert np.abs(mel - gf["f5_mel"]).max() <= 1e-3
    d = np.abs(pcm.astype(np.int32) - gf["f5_pcm"].astype(np.int32))
    assert d.max() <= 3 and (d <= 1).mean() > 0.99
    _, mel1 = f5_engine.f5_synthesize(audio, ids, N, noise, precision=capi.F32, n_steps=1, return_mel=True)
    assert np.abs(mel1 - gf["f5_mel_after_1"]).max() <= 1e-4
    try:
        for prec, cos_floor, snr_floor in ((capi.F16, 0.99999, 55.0), (capi.BF16, 0.9999, 40.0)):
            for chain in (1, 0):
                f5_engine.set_option("dit_chain", chain)
                pcm, mel = f5_engine.f5_synthesize(audio, ids, N, noise, precision=prec, return_mel=True)
                assert np.isfinite(mel).all()
                assert cosine(mel, gf["f5_mel"]) >= cos_floor, (prec, chain, cosine(mel, gf["f5_mel"]))
                assert cosine(mel[:, ref_len:], gf["f5_mel"][:, ref_len:]) >= cos_floor
                assert snr_db(gf["f5_pcm"], pcm) >= snr_floor, (prec, chain, snr_db(gf["f5_pcm"], pcm))
    finally:
        f5_engine.set_option("dit_chain", 1)


def test_f5_config3_fp8_option_vs_reference_golden(f5_engine, gf):
    """Engine option dit_fp8 (1 / 2): ff1 and q|k|v (and ff2) of the fused chain with e4m3 operands (tcgen05 kind::f8f6f4, per-row
    activation scale -- a static one for the hidden activation --, per-output-channel weight scale, fp32 accumulation). An optional lower-fidelity mode: measured on B200 mel cosine 0.99993,
    generated-mel cosine 0.99981, PCM SNR 32.4 dB at N = 1126 (oracle/fp8_study.py predicts 29.7 dB for e4m3 operands in these
    GEMMs); the bars below are what the mode must keep. Switching it off again restores the fp16 numbers."""
    audio, ids, maxd, noise = synth.f5_inputs(int(gf["input_seed"]), int(gf["audio_len"]), int(gf["n_text"]))
    N = int(maxd[0])
    ref_len = int(gf["f5_ref_signal_len"])
    try:
        # level 1: ff1 + q|k|v (32.4 dB measured); level 2: ff2 as well, fed by an e4m3 hidden activation (30.5 dB measured)
        for level, cos_floor, gen_floor, snr_floor in ((1, 0.9998, 0.9995, 28.0), (2, 0.9997, 0.9993, 26.0)):
            f5_engine.set_option("dit_fp8", level)
            pcm, mel = f5_engine.f5_synthesize(audio, ids, N, noise, precision=capi.F16, return_mel=True)
            assert np.isfinite(mel).all()
            assert cosine(mel, gf["f5_mel"]) >= cos_floor, (level, cosine(mel, gf["f5_mel"]))
            assert cosine(mel[:, ref_len:], gf["f5_mel"][:, ref_len:]) >= gen_floor
            assert snr_db(gf["f5_pcm"], pcm) >= snr_floor, (level, snr_db(gf["f5_pcm"], pcm))
            assert snr_db(gf["f5_pcm"], pcm) < 50.0            # the option really changed the arithmetic
    finally:
        f5_engine.set_option("dit_fp8", 0)
    pcm, mel = f5_engine.f5_synthesize(audio, ids, N, noise, precision=capi.F16, return_mel=True)
    assert snr_db(gf["f5_pcm"], pcm) >= 55.0


def test_f5_bigvgan_pipeline_vs_reference_golden(f5_engine, bigvgan_engine, gf):
    """The metric's pipeline through ONE host-buffer C call (b200tts_f5_bigvgan_pipeline): the mel equals the F5 golden, the Vocos
    wav the F5_Decode golden, and the BigVGAN wav equals the BigVGAN session fed that mel (parity is per graph: the reference never
    chains the two). U = 2 copies of the golden utterance: the batched DiT loop sees M = 4504 rows."""
    audio, ids, maxd, noise = synth.f5_inputs(int(gf["input_seed"]), int(gf["audio_len"]), int(gf["n_text"]))
    N, ref_len = int(maxd[0]), int(gf["f5_ref_signal_len"])
    U = 2
    au = np.repeat(audio.reshape(1, -1), U, 0)
    tx = np.repeat(ids.reshape(1, -1), U, 0)
    nz = np.repeat(noise.reshape(1, N, 100), U, 0)
    wav, voc, mel = f5_engine.f5_bigvgan_pipeline(au, tx, N, nz, precision=capi.F16, with_vocos=True, return_mel=True)
    G = N - ref_len
    assert wav.shape == (U, 256 * G + 30) and voc.shape == (U, 256 * (G - 1)) and mel.shape == (U, N, 100)
    for u in range(U):
        assert cosine(mel[u], gf["f5_mel"][0]) >= 0.99999
        assert snr_db(gf["f5_pcm"].reshape(-1), voc[u]) >= 55.0
        direct = bigvgan_engine.bigvgan_run(np.ascontiguousarray(mel[u:u + 1, ref_len:, :].transpose(0, 2, 1)), precision=capi.F16)
        np.testing.assert_array_equal(direct.reshape(-1), wav[u])
    # two copies of one utterance in a batch sit at different offsets of the 256-row blocks / tiles: a row's result must not
    # depend on where it sits (the fused LayerNorm spells out its rounding steps for exactly this reason, dit_chain.cu)
    np.testing.assert_array_equal(mel[0], mel[1])
    np.testing.assert_array_equal(wav[0], wav[1])
    wav1 = f5_engine.f5_bigvgan_pipeline(au[:1], tx[:1], N, nz[:1], precision=capi.F16)
    # batch of two vs one: the fused chain splits a row's 1024 columns over 4 instead of 8 CTA pairs, so the LayerNorm partial
    # sums group differently (last-bit differences in the statistics; the vocoder amplifies them): 59.3 dB measured
    assert snr_db(wav[0], wav1[0]) >= 55.0


def test_f5_config3_batch_equals_single(f5_engine):
    """configs[2] shape (6 s reference, 150 text ids, N = 1126), 3 Euler steps: U = 4 batched == 4 single utterances."""
    U, L, n_text, steps = 4, 144000, 150, 3
    ins = [synth.f5_inputs(70 + i, L, n_text) for i in range(U)]
    N = int(ins[0][2][0])
    assert N == 1126
    ns = 256 * (N - (L // 256 + 1) - 1)
    audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
    ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
    noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
    pcm_b = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
    mel_b = torch.zeros((U, N, FCFG.n_mels), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    f5_engine.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm_b.data_ptr(),
                                         precision=capi.BF16, n_steps=steps, mel_ptr=mel_b.data_ptr())
    f5_engine.synchronize()
    assert bool(torch.isfinite(mel_b).all())
    for u in (0, 3):
        pcm_1 = torch.zeros((ns,), dtype=torch.int16, device="cuda")
        mel_1 = torch.zeros((N, FCFG.n_mels), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        f5_engine.f5_synthesize_device(audio[u].data_ptr(), L, ids[u].data_ptr(), n_text, N, noise[u].data_ptr(), pcm_1.data_ptr(),
                                       precision=capi.BF16, n_steps=steps, mel_ptr=mel_1.data_ptr())
        f5_engine.synchronize()
        np.testing.assert_allclose(mel_b[u].cpu().numpy(), mel_1.cpu().numpy(), rtol=0, atol=1e-5)
        assert np.abs(pcm_b[u].cpu().numpy().astype(np.int32) - pcm_1.cpu().numpy().astype(np.int32)).max() <= 1
    # the reference region (first 563 frames) is driven by the same conditioning as the generated one: both move off the noise
    moved = (mel_b - noise.view(U, N, FCFG.n_mels)).abs().mean(dim=(1, 2))
    assert bool((moved > 1e-3).all())


def test_f5_two_phase_chain_schedule_equals_single(f5_engine):
    """Nine uniform config-3 utterances = 80 row blocks: the fused chain walks 74 of them in phase 0 (one CTA pair each) and the last
    6 in phase 1 (a team of 8 pairs each, other weight maps) -- the schedule configs[3] runs with (593 = 8 x 74 + 1 blocks). The first
    utterance (phase 0) and the last one (its rows lie in the phase-1 blocks) must equal their single runs (team of 8, one phase)."""
    U, L, n_text, steps = 9, 144000, 150, 2
    plan = (ctypes.c_int * 5)()
    assert capi.load_library().b200tts_debug_chain_plan((2 * U * 1126 + 255) // 256, 74, plan) == 0
    assert list(plan) == [1, 74, 74, 6, 8]
    ins = [synth.f5_inputs(170 + i, L, n_text) for i in range(U)]
    N = int(ins[0][2][0])
    ns = 256 * (N - (L // 256 + 1) - 1)
    audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
    ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
    noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
    pcm_b = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
    mel_b = torch.zeros((U, N, FCFG.n_mels), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for level in (0, 2):                                     # fp16 operands, and the e4m3 option (its phase-1 weight maps differ too)
        try:
            f5_engine.set_option("dit_fp8", level)
            f5_engine.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm_b.data_ptr(),
                                                 precision=capi.F16, n_steps=steps, mel_ptr=mel_b.data_ptr())
            f5_engine.synchronize()
            assert bool(torch.isfinite(mel_b).all())
            for u in (0, U - 1):
                pcm_1 = torch.zeros((ns,), dtype=torch.int16, device="cuda")
                mel_1 = torch.zeros((N, FCFG.n_mels), dtype=torch.float32, device="cuda")
                torch.cuda.synchronize()
                f5_engine.f5_synthesize_device(audio[u].data_ptr(), L, ids[u].data_ptr(), n_text, N, noise[u].data_ptr(), pcm_1.data_ptr(),
                                               precision=capi.F16, n_steps=steps, mel_ptr=mel_1.data_ptr())
                f5_engine.synchronize()
                np.testing.assert_allclose(mel_b[u].cpu().numpy(), mel_1.cpu().numpy(), rtol=0, atol=1e-5)
        finally:
            f5_engine.set_option("dit_fp8", 0)


def test_f5_config3_full_run_is_deterministic(f5_engine):
    """All 31 Euler steps at N = 1126, twice (eager, then graph replay): identical PCM, finite mel, plausible level."""
    L, n_text = 144000, 150
    audio, ids, maxd, noise = synth.f5_inputs(1, L, n_text)
    N = int(maxd[0])
    a = f5_engine.f5_synthesize(audio, ids, N, noise, precision=capi.BF16)
    b = f5_engine.f5_synthesize(audio, ids, N, noise, precision=capi.BF16)
    c = f5_engine.f5_synthesize(audio, ids, N, noise, precision=capi.BF16)
    assert a.shape[-1] == 256 * (N - (L // 256 + 1) - 1) == 143872
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, c)
    assert np.abs(a.astype(np.int32)).max() > 10

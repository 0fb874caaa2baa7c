"""Round-2 GPU probe: parity numbers and timings of the new paths, one section per process (a device trap in one section
must not poison the next). Prints one JSON line per measurement; tools/r02_probe.sh runs the sections under `timeout`.

    python tools/r02_probe.py <section> [...]

Sections: f5_small, f5_full, f5_time, bigvgan, pipeline"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import b200tts  # noqa: E402,F401
from b200tts import capi, config, synth, weights  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PREC = {"f32": capi.F32, "bf16": capi.BF16, "f16": capi.F16}


def out(**kw):
    print(json.dumps(kw), flush=True)


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return float(10.0 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30)))


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def f5_engine():
    eng = capi.Engine(0)
    cfg = config.F5
    dsd = synth.f5_dit_state(4321)
    eng.load_state("dit", weights.dit_engine_tensors(dsd, cfg))
    eng.load_state("vocos", weights.vocos_engine_tensors(synth.vocos_state(2468), cfg))
    eng.load_state("f5", weights.f5_export_constants(dsd, cfg))
    eng.f5_build()
    return eng


def sec_f5_small():
    g = dict(np.load(os.path.join(GOLD, "f5_ref.npz")))
    eng = f5_engine()
    audio, text_ids, maxd, noise = synth.f5_inputs(int(g["input_seed"]), int(g["audio_len"]), int(g["n_text"]))
    N = int(maxd[0])
    res = {}
    for prec in ("bf16", "f16"):
        for chain in (0, 1):
            eng.set_option("dit_chain", chain)
            pcm, mel = eng.f5_synthesize(audio, text_ids, N, noise, precision=PREC[prec], return_mel=True)
            res[(prec, chain)] = (pcm, mel)
            out(section="f5_small", prec=prec, chain=chain, N=N, mel_cos=cosine(mel, g["noise_after_31"]),
                mel_maxabs=float(np.abs(mel - g["noise_after_31"]).max()), pcm_snr=snr_db(g["pcm"], pcm), finite=bool(np.isfinite(mel).all()))
        a, b = res[(prec, 0)], res[(prec, 1)]
        out(section="f5_small", prec=prec, what="chain vs unfused", mel_maxabs=float(np.abs(a[1] - b[1]).max()), pcm_snr=snr_db(a[0], b[0]))
    # one step only: tighter comparison of the two code paths
    for prec in ("bf16", "f16"):
        xs = []
        for chain in (0, 1):
            eng.set_option("dit_chain", chain)
            x1, _ = eng.f5_transformer(noise, g["rope_cos_row"], g["rope_sin_row"], g["cat_mel_text"], g["cat_mel_text_drop"], 0,
                                       n_steps=1, precision=PREC[prec])
            xs.append(x1)
        dt0 = float(g["delta_t"][0])
        p0, p1, pr = (xs[0] - noise) / dt0, (xs[1] - noise) / dt0, (g["noise_after_1"] - noise) / dt0
        out(section="f5_small", prec=prec, what="one step prediction", cos_unfused=cosine(p0, pr), cos_chain=cosine(p1, pr),
            cos_chain_vs_unfused=cosine(p0, p1), maxabs_chain_vs_unfused=float(np.abs(p0 - p1).max()), pred_rms=float(np.sqrt((pr ** 2).mean())))


def sec_f5_full():
    g = dict(np.load(os.path.join(GOLD, "fullsize_ref.npz")))
    eng = f5_engine()
    audio, text_ids, maxd, noise = synth.f5_inputs(int(g["input_seed"]), int(g["audio_len"]), int(g["n_text"]))
    N = int(maxd[0])
    t0 = time.time()
    pcm, mel = eng.f5_synthesize(audio, text_ids, N, noise, precision=capi.F32, return_mel=True)
    d = np.abs(pcm.astype(np.int32) - g["f5_pcm"].astype(np.int32))
    out(section="f5_full", prec="f32", N=N, mel_maxabs=float(np.abs(mel - g["f5_mel"]).max()), pcm_max_lsb=int(d.max()),
        pcm_le1=float((d <= 1).mean()), pcm_snr=snr_db(g["f5_pcm"], pcm), wall_s=time.time() - t0)
    for prec in ("bf16", "f16"):
        for chain in (0, 1):
            eng.set_option("dit_chain", chain)
            pcm, mel = eng.f5_synthesize(audio, text_ids, N, noise, precision=PREC[prec], return_mel=True)
            out(section="f5_full", prec=prec, chain=chain, mel_cos=cosine(mel, g["f5_mel"]), mel_maxabs=float(np.abs(mel - g["f5_mel"]).max()),
                gen_mel_cos=cosine(mel[:, int(g["f5_ref_signal_len"]):], g["f5_mel"][:, int(g["f5_ref_signal_len"]):]),
                pcm_snr=snr_db(g["f5_pcm"], pcm), finite=bool(np.isfinite(mel).all()))
            pcm1, mel1 = eng.f5_synthesize(audio, text_ids, N, noise, precision=PREC[prec], n_steps=1, return_mel=True)
            out(section="f5_full", prec=prec, chain=chain, what="after 1 step", mel_maxabs=float(np.abs(mel1 - g["f5_mel_after_1"]).max()),
                pred_cos=cosine(mel1 - noise, g["f5_mel_after_1"] - noise))


def sec_f5_time():
    import torch
    eng = f5_engine()
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    L, n_text = 144000, 150
    cfg = config.F5
    for U in (1, 8):
        ins = [synth.f5_inputs(1000 + i, L, n_text) for i in range(U)]
        N = int(ins[0][2][0])
        ns = 256 * (N - (L // 256 + 1) - 1)
        audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
        ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
        noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
        pcm = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
        for prec in ("bf16", "f16"):
            for chain in (0, 1):
                eng.set_option("dit_chain", chain)

                def run():
                    eng.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm.data_ptr(),
                                                   precision=PREC[prec])
                with torch.cuda.stream(stream):
                    for _ in range(3):
                        run()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 5 if U == 1 else 2
                    e0.record(stream)
                    for _ in range(reps):
                        run()
                    e1.record(stream)
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                out(section="f5_time", U=U, prec=prec, chain=chain, ms_per_call=ms, ms_per_utt=ms / U,
                    dit_tflops=U * 33.5e12 / (ms / 1e3) / 1e12)
                if U == 1:
                    eng.profile_begin()
                    with torch.cuda.stream(stream):
                        run()
                    prof = eng.profile_end()
                    out(section="f5_time", U=U, prec=prec, chain=chain, profile_ms={k: round(v["ms"], 3) for k, v in prof.items() if v["ms"] > 0.3})


def sec_bigvgan():
    import torch
    g = dict(np.load(os.path.join(GOLD, "fullsize_ref.npz")))
    eng = capi.Engine(0)
    eng.load_state("bigvgan", weights.bigvgan_engine_tensors(synth.bigvgan_state(int(g["vgan_seed"]))))
    eng.bigvgan_build()
    mel = synth.bigvgan_mel(int(g["vgan_mel_seed"]), 1, int(g["vgan_T"]))
    want = g["vgan_pcm"].astype(np.int32)
    got = eng.bigvgan_run(mel, precision=capi.F32).astype(np.int32)
    d = np.abs(got - want)
    out(section="bigvgan", prec="f32", max_lsb=int(d.max()), le1=float((d <= 1).mean()), snr=snr_db(want, got))
    mel8 = synth.bigvgan_mel(100, 8, 512)
    mel8[3] = mel[0]
    for prec in ("bf16", "f16"):
        got = eng.bigvgan_run(mel, precision=PREC[prec])
        got8 = eng.bigvgan_run(mel8, precision=PREC[prec])
        out(section="bigvgan", prec=prec, snr_b1=snr_db(want, got), snr_in_batch8=snr_db(want, got8[3:4]),
            batch_item_equal=bool(np.array_equal(got[0], got8[3])))
        stream = torch.cuda.Stream()
        eng.set_stream(stream.cuda_stream)
        md = torch.from_numpy(mel8).cuda()
        pd = torch.zeros((8, 1, config.BIGVGAN.out_samples(512)), dtype=torch.int16, device="cuda")
        with torch.cuda.stream(stream):
            for _ in range(3):
                eng.bigvgan_run_device(md.data_ptr(), 8, 512, pd.data_ptr(), precision=PREC[prec])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                eng.bigvgan_run_device(md.data_ptr(), 8, 512, pd.data_ptr(), precision=PREC[prec])
            e1.record(stream)
            torch.cuda.synchronize()
        out(section="bigvgan", prec=prec, ms_per_step_b8=e0.elapsed_time(e1) / 10)
        eng.set_stream(0)


def sec_bvg_branches():
    """BigVGAN with the three resblock branches of a stage serial (0) or concurrent (1): outputs must be bit-identical; time per step
    for the configs[1] batch (8, 100, 512) and for one config-3 utterance's generated mel (1, 100, 563)."""
    import torch
    eng = capi.Engine(0)
    eng.load_state("bigvgan", weights.bigvgan_engine_tensors(synth.bigvgan_state(1234)))
    eng.bigvgan_build()
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    res = {}
    for B, T in ((8, 512), (1, 563)):
        mel = synth.bigvgan_mel(100, B, T)
        md = torch.from_numpy(mel).cuda()
        for br in (0, 1):
            eng.set_option("bigvgan_branches", br)
            pd = torch.zeros((B, 1, config.BIGVGAN.out_samples(T)), dtype=torch.int16, device="cuda")
            with torch.cuda.stream(stream):
                for _ in range(3):
                    eng.bigvgan_run_device(md.data_ptr(), B, T, pd.data_ptr(), precision=capi.F16)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(10):
                    eng.bigvgan_run_device(md.data_ptr(), B, T, pd.data_ptr(), precision=capi.F16)
                e1.record(stream)
                torch.cuda.synchronize()
            res[(B, T, br)] = pd.cpu().numpy().copy()
            out(section="bvg_branches", B=B, T=T, branches=br, ms_per_step=e0.elapsed_time(e1) / 10)
        out(section="bvg_branches", B=B, T=T, bit_identical=bool(np.array_equal(res[(B, T, 0)], res[(B, T, 1)])))


def sec_f5_fp8():
    """Optional e4m3 mode of ff1 / q|k|v (engine option dit_fp8): parity against the reference goldens at N = 130 and N = 1126, and
    time per call for one and eight config-3 utterances, fp8 off / on."""
    import torch
    eng = f5_engine()
    for gname, mel_key, pcm_key, ref_key in (("f5_ref.npz", "noise_after_31", "pcm", "ref_signal_len"), ("fullsize_ref.npz", "f5_mel", "f5_pcm", "f5_ref_signal_len")):
        g = dict(np.load(os.path.join(GOLD, gname)))
        audio, text_ids, maxd, noise = synth.f5_inputs(int(g["input_seed"]), int(g["audio_len"]), int(g["n_text"]))
        N = int(maxd[0])
        for fp8 in (0, 1, 2):
            eng.set_option("dit_fp8", fp8)
            pcm, mel = eng.f5_synthesize(audio, text_ids, N, noise, precision=capi.F16, return_mel=True)
            r = int(g[ref_key])
            out(section="f5_fp8", N=N, fp8=fp8, mel_cos=cosine(mel, g[mel_key]), gen_mel_cos=cosine(mel[:, r:], g[mel_key][:, r:]),
                mel_maxabs=float(np.abs(mel - g[mel_key]).max()), pcm_snr=snr_db(g[pcm_key], pcm), finite=bool(np.isfinite(mel).all()))
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    L, n_text = 144000, 150
    for U in (1, 8):
        ins = [synth.f5_inputs(1000 + i, L, n_text) for i in range(U)]
        N = int(ins[0][2][0])
        ns = 256 * (N - (L // 256 + 1) - 1)
        audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
        ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
        noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
        pcm = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
        for fp8 in (0, 1, 2):
            eng.set_option("dit_fp8", fp8)

            def run():
                eng.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm.data_ptr(), precision=capi.F16)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5 if U == 1 else 2
                e0.record(stream)
                for _ in range(reps):
                    run()
                e1.record(stream)
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            eng.profile_begin()
            with torch.cuda.stream(stream):
                run()
            prof = eng.profile_end()
            out(section="f5_fp8", U=U, fp8=fp8, ms_per_call=ms, ms_per_utt=ms / U, chain_ms=round(prof["f5.chain"]["ms"], 3),
                attention_ms=round(prof["f5.attention"]["ms"], 3))


def sec_attn_time():
    """fp16, fused chain: time per call and the event-timed share of attention / chain, one and eight config-3 utterances."""
    import torch
    eng = f5_engine()
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    L, n_text = 144000, 150
    for U in (1, 8):
        ins = [synth.f5_inputs(1000 + i, L, n_text) for i in range(U)]
        N = int(ins[0][2][0])
        ns = 256 * (N - (L // 256 + 1) - 1)
        audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
        ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
        noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
        pcm = torch.zeros((U, ns), dtype=torch.int16, device="cuda")

        def run():
            eng.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm.data_ptr(), precision=capi.F16)
        with torch.cuda.stream(stream):
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if U == 1 else 2
            e0.record(stream)
            for _ in range(reps):
                run()
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        eng.profile_begin()
        with torch.cuda.stream(stream):
            run()
        prof = eng.profile_end()
        out(section="attn_time", U=U, poly=os.environ.get("B200TTS_ATTN_POLY", "default"), ms_per_call=ms, ms_per_utt=ms / U,
            attention_ms=round(prof["f5.attention"]["ms"], 3), chain_ms=round(prof["f5.chain"]["ms"], 3))


def sec_twins():
    """Two copies of one utterance in a batch: are their mels bit-identical? (rows at different tile offsets)"""
    import torch
    eng = f5_engine()
    L, n_text = int(os.environ.get("TWIN_L", "144000")), 150
    a, t, maxd, nz = synth.f5_inputs(1, L, n_text)
    N = int(os.environ.get("TWIN_N", int(maxd[0])))
    nz = np.random.default_rng(5).standard_normal((1, N, 100), dtype=np.float32)
    U = 2
    ns = 256 * (N - (L // 256 + 1) - 1)
    audio = torch.from_numpy(np.repeat(a.reshape(1, -1), U, 0)).cuda()
    ids = torch.from_numpy(np.repeat(t.reshape(1, -1), U, 0)).cuda()
    noise = torch.from_numpy(np.repeat(nz.reshape(1, -1), U, 0)).cuda()
    pcm = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
    mel = torch.zeros((U, N, 100), dtype=torch.float32, device="cuda")
    for prec in ("f16",):
        for chain in (0, 1):
            eng.set_option("dit_chain", chain)
            for steps in (1,):
                eng.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm.data_ptr(),
                                               precision=PREC[prec], n_steps=steps, mel_ptr=mel.data_ptr())
                eng.synchronize()
                m = mel.cpu().numpy()
                d = np.abs(m[0] - m[1])
                rows = np.nonzero(d.max(axis=1))[0]
                out(section="twins", N=N, team=os.environ.get("B200TTS_CHAIN_TEAM", "auto"), prec=prec, chain=chain, steps=steps, mel_maxabs_between_twins=float(d.max()), n_rows_differ=int(rows.size),
                    first_rows=rows[:8].tolist(), pcm_equal=bool((pcm[0] == pcm[1]).all().item()))


if __name__ == "__main__":
    for name in sys.argv[1:]:
        t0 = time.time()
        globals()["sec_" + name]()
        out(section=name, done=True, wall_s=time.time() - t0)

#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the B200 TTS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload pipeline|f5|bigvgan|...]

One "step" = one pass of the hot path over one batch of synthetic input.
  config4  (default) BASELINE.json's metric configuration ("F5-TTS NFE=32 + BigVGAN 24 kHz"), configs[3]: a batch of 64 synthetic
           utterances (references 4-8 s, N in [752, 1502]) dealt longest-first to the N GPUs; a rank runs its share as ONE ragged
           batch through b200tts_f5_bigvgan_pipeline_ragged -- graph A per utterance, one 31-step DiT loop over all its rows,
           BigVGAN on every utterance's generated frames. STRONG scaling: the 64 utterances are the job whatever N is.
           `value`: inputs resident in HBM (device-pointer entry); `e2e`: the host-buffer C call (pinned host buffers, H2D / D2H
           inside the timed region).
  pipeline the same pipeline on U = 8 utterances of config-3 shape (6 s reference) per GPU, weak scaling.
  f5       configs[2]: one utterance per step (latency): preprocess + DiT loop + Vocos/ISTFT decode.
  bigvgan  configs[1]: BigVGAN-v2 24khz_100band_256x, mels (8,100,512) -> int16 PCM.
The default run (one GPU) attaches short `pipeline`, `f5` and `bigvgan` measurements under those keys. Utterances shard with no
data-path collective; NCCL is used once, to broadcast the weights from rank 0 at load. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def ncu_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the named kernel from the committed
    `ncu --set full` capture (profiles/r02/traffic.json, then profiles/r01/traffic.json, say which launch of which capture); None if not captured."""
    for rnd in ("r02", "r01"):                       # the latest capture of a kernel wins
        p = os.path.join(ROOT, "profiles", rnd, "traffic.json")
        if os.path.exists(p):
            e = json.load(open(p)).get(key)
            if e is not None:
                return e["bytes_per_launch"]
    return None


def profiled(eng, torch, stream, fn):
    """One extra, untimed pass with a CUDA-event pair around every launch (engine profiler) -> {tag: {"ms", "launches"}}. BigVGAN's
    concurrent resblock branches are serialised for this pass only: with three streams in flight a kernel's event pair also spans
    the time it waits for SMs held by the other branches, and the per-kernel sums (roofline.achieved, share_of_step) would be
    inflated. The timed steps always run the default (concurrent) schedule."""
    eng.set_option("bigvgan_branches", 0)
    try:
        eng.profile_begin()
        with torch.cuda.stream(stream):
            fn()
        return eng.profile_end()
    finally:
        eng.set_option("bigvgan_branches", 1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# work models (DESIGN.md "Work model")
# ------------------------------------------------------------------------------------------------------
def bigvgan_work(cfg, B, T):
    """Algorithmic FLOPs / elements of one BigVGAN pass over B mels of T frames."""
    L, C = T, cfg.upsample_initial_channel
    flops_res = flops_up = 0.0
    aa_elems = 0
    for u, k in zip(cfg.upsample_rates, cfg.upsample_kernel_sizes):
        flops_up += 2.0 * C * (C // 2) * k * L      # every input sample meets k taps of C/2 outputs
        C //= 2
        L *= u
        for rk in cfg.resblock_kernel_sizes:
            flops_res += 6 * 2.0 * C * C * rk * L
            aa_elems += 6 * C * L
    flops_pre = 2.0 * cfg.num_mels * cfg.upsample_initial_channel * 7 * T
    flops_post = 2.0 * C * 7 * (L + 30)
    aa_elems += C * (L + 30)
    return {"frames": B * T, "flops_resconv": B * flops_res, "flops_total": B * (flops_res + flops_up + flops_pre + flops_post),
            "aa_elems": B * aa_elems, "audio_s": B * cfg.out_samples(T) / cfg.sample_rate}


def bigvgan_stage_table(cfg, B, T, profile_ms, pk, bytes_per_elem=2):
    """Per upsampling stage: the resblock work against BOTH rooflines SURVEY.md 8(d) names. FLOPs = the 18 convs of the stage;
    bytes = the survey's layer-fused model M1, 49 * C * L elements per stage (every conv reads its input and writes its output once,
    the anti-aliased activations ride along, + the residual / accumulate reads); time = the stage's conv launches plus its
    activation launches (`profile_ms`: the per-kernel event times of one serialised pass). `bound` = whichever floor is higher."""
    rows, L, C = [], T, cfg.upsample_initial_channel
    for i, u in enumerate(cfg.upsample_rates):
        C //= 2
        L *= u
        ms = profile_ms.get(f"bigvgan.resconv.s{i}", 0.0) + profile_ms.get(f"bigvgan.aa_snake.s{i}", 0.0)
        if ms <= 0:
            continue
        flops = B * sum(6 * 2.0 * C * C * rk * L for rk in cfg.resblock_kernel_sizes)
        m1 = B * 49.0 * C * L * bytes_per_elem
        t_tensor, t_hbm = flops / (pk["bf16_tflops_sustained"] * 1e12) * 1e3, m1 / (pk["hbm_gbs"] * 1e9) * 1e3
        rows.append({"stage": i, "channels": C, "samples": L, "ms": round(ms, 4), "conv_ms": round(profile_ms.get(f"bigvgan.resconv.s{i}", 0.0), 4),
                     "activation_ms": round(profile_ms.get(f"bigvgan.aa_snake.s{i}", 0.0), 4),
                     "tensor_frac": round(t_tensor / ms, 4), "hbm_frac_m1": round(t_hbm / ms, 4),
                     "bound": "tensor" if t_tensor >= t_hbm else "hbm", "floor_ms": round(max(t_tensor, t_hbm), 4)})
    return rows


def f5_work(cfg, N, ref_len):
    """Algorithmic FLOPs of the DiT loop for one utterance of N frames (both CFG rows)."""
    D, FF = cfg.dim, cfg.dim * cfg.ff_mult
    per_tok_layer_gemm = 2.0 * (3 * D * D + D * D + 2 * D * FF)         # qkv, out, ff1, ff2
    per_tok_layer_attn = 4.0 * D * N                                    # QK^T and PV over all heads
    gc = D // cfg.convpos_groups
    per_tok_embed = 2.0 * cfg.n_mels * D + 2 * 2.0 * gc * cfg.convpos_kernel * D + 2.0 * D * cfg.n_mels
    per_tok = cfg.depth * (per_tok_layer_gemm + per_tok_layer_attn) + per_tok_embed
    steps = cfg.nfe - 1
    G = N - ref_len
    # the fused row-block chain (dit_chain.cu) runs out-proj + ff1 + ff2 of every block and q|k|v of blocks 1..depth-1
    per_tok_chain = cfg.depth * 2.0 * (D * D + 2 * D * FF) + (cfg.depth - 1) * 2.0 * 3 * D * D
    return {"frames": G, "flops_step": 2 * N * per_tok, "flops_gemm_step": 2 * N * cfg.depth * per_tok_layer_gemm,
            "flops_chain_step": 2 * N * per_tok_chain,
            "flops_attn_step": 2 * N * cfg.depth * per_tok_layer_attn, "flops_total": steps * 2 * N * per_tok,
            "steps": steps, "audio_s": cfg.hop * (G - 1) / cfg.sample_rate}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path (oracle restatement of its PyTorch modules; ORT is not installable
# offline -- BASELINE.md section 2), all host threads, bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------
def cpu_bigvgan(T, reps, cores):
    import torch
    import b200tts  # noqa: F401
    from b200tts import config, synth
    from oracle import bigvgan_ref
    torch.set_num_threads(cores)
    cfg = config.BIGVGAN
    sd = synth.bigvgan_state(1234)
    mel = synth.bigvgan_mel(100, 1, T)
    bigvgan_ref.bigvgan_pcm(mel[:, :, :32], sd, cfg)
    t0 = time.perf_counter()
    for _ in range(reps):
        bigvgan_ref.bigvgan_pcm(mel, sd, cfg)
    return (time.perf_counter() - t0) / reps


def cpu_f5(N_audio, n_text, dit_steps, cores):
    """Times graph A, `dit_steps` DiT steps and graph C of the oracle; returns seconds per part."""
    import torch
    import b200tts  # noqa: F401
    from b200tts import config, synth
    from oracle import f5_ref as R
    torch.set_num_threads(cores)
    cfg = config.F5
    dsd, vsd = synth.f5_dit_state(4321), synth.vocos_state(2468)
    audio, text_ids, maxd, noise = synth.f5_inputs(1, N_audio, n_text)
    with torch.inference_mode():
        sd = R.prescale_qk(dsd, cfg)
        tables = R.time_tables(sd, cfg)
        t0 = time.perf_counter()
        x, cq, sq, _, _, cond, cond_drop, ref_len = R.f5_preprocess(audio, text_ids, maxd, sd, cfg, noise)
        t_pre = time.perf_counter() - t0
        cos, sin = cq[0, 0], sq[0, 0]
        ts = 0
        t0 = time.perf_counter()
        for _ in range(dit_steps):
            x, ts = R.f5_transformer_step(sd, x, cond, cond_drop, ts, tables, cfg, cos, sin)
        t_step = (time.perf_counter() - t0) / max(dit_steps, 1)
        fv = R.fold_vocos(vsd, cfg)
        t0 = time.perf_counter()
        R.f5_decode(x, ref_len, fv, cfg)
        t_dec = time.perf_counter() - t0
    return {"pre_s": t_pre, "step_s": t_step, "decode_s": t_dec, "N": int(maxd[0]), "ref_len": int(ref_len)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import b200tts  # noqa: F401
    from b200tts import config
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 3))
    if args.workload == "bigvgan":
        cfg = config.BIGVGAN
        T = args.frames
        dt = cpu_bigvgan(T, steps, cores)
        v = T / dt
        line = {"metric": "mel_frames_per_s", "value": v, "unit": "mel-frames/s", "ms_per_step": 1e3 * dt, "dtype": "f32",
                "config": {"workload": f"BigVGAN-v2 24khz_100band_256x, mels ({args.batch},100,{T}) [BASELINE.json configs[1]]; "
                                       f"each reference step = 1 mel of that batch", "parallelism": "cpu"},
                "rtf": dt / (cfg.out_samples(T) / cfg.sample_rate),
                "cpu_baseline": {"value": v, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                                 "sample": f"{steps} x 1 mel (1,100,{T}), torch-CPU fp32 eager restatement of the reference "
                                           "modules (stand-in for ORT CPUExecutionProvider, not installable offline)"}}
    elif args.workload in ("indextts_gpt", "indextts"):     # the acoustic half dominates the CPU time of a sentence
        import torch as _t
        from b200tts import synth
        from oracle import indextts_gpt_ref
        _t.set_num_threads(cores)
        cfg = config.INDEXTTS_GPT
        sd = synth.igpt_state(555)
        cds, tid = synth.igpt_inputs(900, args.gpt_text, cfg)
        t0 = time.perf_counter()
        indextts_gpt_ref.generate(cds, tid, sd, cfg, max_new=1)
        t_pre = time.perf_counter() - t0
        n_cpu = 1 + 8 * steps
        t0 = time.perf_counter()
        indextts_gpt_ref.generate(cds, tid, sd, cfg, max_new=n_cpu)
        per_tok = max(time.perf_counter() - t0 - t_pre, 1e-9) / (n_cpu - 1)
        total = t_pre + (args.new_tokens - 1) * per_tok
        v = args.new_tokens / total
        line = {"metric": "mel_tokens_per_s", "value": v, "unit": "tokens/s", "ms_per_step": 1e3 * total, "dtype": "f32",
                "config": {"workload": f"IndexTTS GPT-2 acoustic model, one sentence: prefill of {cfg.cond_rows + args.gpt_text + 3} rows + "
                                       f"{args.new_tokens - 1} decode calls [acoustic half of BASELINE.json configs[4]]", "parallelism": "cpu"},
                "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port",
                                 "sample": f"prefill + {n_cpu - 1} decode calls (extrapolated to {args.new_tokens}); torch-CPU fp32 eager "
                                           "restatement of graphs B-E", "prefill_s": t_pre, "s_per_decode_call": per_tok}}
    else:
        cfg = config.F5
        r = cpu_f5(args.audio_len, args.n_text, steps, cores)
        G = r["N"] - r["ref_len"]
        total = r["pre_s"] + (cfg.nfe - 1) * r["step_s"]
        if args.workload in ("pipeline", "config4"):
            r["bigvgan_s"] = cpu_bigvgan(G, 1, cores)
            total += r["bigvgan_s"]
        else:
            total += r["decode_s"]
        v = G / total
        line = {"metric": "mel_frames_per_s", "value": v, "unit": "mel-frames/s", "ms_per_step": 1e3 * total, "dtype": "f32",
                "config": {"workload": f"F5-TTS NFE={cfg.nfe} N={r['N']} (ref {r['ref_len']} frames)"
                                       + (" + BigVGAN on the generated frames [BASELINE.json configs[3] per-GPU share; one of its utterances per reference step]"
                                          if args.workload in ("pipeline", "config4") else " + Vocos/ISTFT [BASELINE.json configs[2]]; one utterance"),
                           "parallelism": "cpu"},
                "rtf": total / (cfg.hop * (G - 1) / cfg.sample_rate),
                "cpu_baseline": {"value": v, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                                 "sample": f"graph A once, {steps} of {cfg.nfe - 1} DiT steps (x{cfg.nfe - 1} extrapolated), "
                                           + ("BigVGAN on the generated frames once" if args.workload in ("pipeline", "config4") else "graph C once")
                                           + "; torch-CPU fp32 eager restatement of the reference modules (stand-in for ORT "
                                             "CPUExecutionProvider, not installable offline)", "parts_s": r}}
    line.update({"impl": "reference", "n_gpus": world, "steps": steps, "warmup": 1, "higher_is_better": True, "scaling": "weak",
                 "vs_baseline": None, "data": "synthetic",
                 "e2e": {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
class Harness:
    def __init__(self, torch, dist, stream, world):
        self.torch, self.dist, self.stream, self.world = torch, dist, stream, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        with torch.cuda.stream(self.stream):
            ev0.record(self.stream)
            for _ in range(steps):
                fn()
            ev1.record(self.stream)
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def bench_bigvgan(args, H, eng, rank, B, T, prec, steps, warmup, sampler=None):
    torch = H.torch
    from b200tts import capi, config, synth
    cfg = config.BIGVGAN
    n_out = cfg.out_samples(T)
    work = bigvgan_work(cfg, B, T)
    mel_host = torch.from_numpy(synth.bigvgan_mel(100 + rank, B, T)).pin_memory()
    pcm_host = torch.empty((B, 1, n_out), dtype=torch.int16).pin_memory()
    mel_dev = mel_host.cuda(non_blocking=True)
    pcm_dev = torch.empty((B, 1, n_out), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()

    def step_device():
        eng.bigvgan_run_device(mel_dev.data_ptr(), B, T, pcm_dev.data_ptr(), precision=prec)

    mel_np, pcm_np = mel_host.numpy(), pcm_host.numpy()

    def step_e2e():          # the reference-facing call: host (pinned) mel in, host (pinned) PCM out, copies inside b200tts_bigvgan_run
        eng.bigvgan_run(mel_np, precision=prec, out=pcm_np)

    with torch.cuda.stream(H.stream):
        for _ in range(warmup):
            step_device()
            step_e2e()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    ms = H.timed(step_device, steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_e2e = H.timed(step_e2e, steps)
    prof = profiled(eng, torch, H.stream, step_device)
    pk = peaks()
    frames = work["frames"] * H.world * steps
    res = {
        "value": frames / (ms / 1e3), "ms_per_step": ms / steps, "rtf": (ms / 1e3 / steps) / work["audio_s"],
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "mel-frames/s", "h2d_bytes_per_step": int(mel_host.numel() * 4),
                "d2h_bytes_per_step": int(pcm_host.numel() * 2), "ms_per_step": ms_e2e / steps,
                "rtf": (ms_e2e / 1e3 / steps) / work["audio_s"]},
        "gpu_launches": int(launches), "clocks": clocks,
        "profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
        "workload": f"BigVGAN-v2 24khz_100band_256x, mels ({B},100,{T}) per GPU -> int16 PCM ({B},1,{n_out}) [BASELINE.json configs[1]]",
    }
    def _sum(prefix):
        sel = [v for k, v in prof.items() if k.startswith(prefix)]
        return {"launches": sum(v["launches"] for v in sel), "ms": sum(v["ms"] for v in sel)}
    conv, aa = _sum("bigvgan.resconv"), _sum("bigvgan.aa_snake.")
    if conv["ms"] > 0:
        ach = work["flops_resconv"] / (conv["ms"] / 1e3) / 1e12
        kern = "rowgemm_f32_kernel (SIMT parity engine)" if prec == capi.F32 else "rowgemm_tc3 / tc2sm kernels (108 resblock convs)"
        res["roofline"] = {"bound": "tensor", "kernel": kern, "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": ach / pk["bf16_tflops_sustained"],
                           "traffic": ncu_traffic("bigvgan.resconv") if prec != capi.F32 else None,
                           "peak_source": pk["source"] + " (sustained cuBLAS bf16: kernel timed inside a long step)",
                           "avg_launch_ms": conv["ms"] / max(conv["launches"], 1),
                           "share_of_step": conv["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9)}
    if prec != capi.F32:
        res["stages"] = bigvgan_stage_table(cfg, B, T, res["profile_ms"], pk)
    if aa["ms"] > 0:
        per_elem = 4.0 if prec != capi.F32 else 8.0          # 16 bit in + 16 bit out on the fast path
        gbs = work["aa_elems"] * per_elem / (aa["ms"] / 1e3) / 1e9
        res["roofline_hbm"] = {"bound": "hbm", "kernel": "aa_snake_kernel", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": gbs / pk["hbm_gbs"], "traffic": ncu_traffic("bigvgan.aa_snake") if prec != capi.F32 else None,
                               "avg_launch_ms": aa["ms"] / max(aa["launches"], 1)}
    return res


def bench_f5(args, H, eng, rank, prec, steps, warmup, with_vocoder=False, U=1, sampler=None):
    """U utterances per step. with_vocoder: the metric's pipeline (A, batched 31-step DiT loop, BigVGAN on the generated
    frames) in one call; else A + 31 x B + C (Vocos / ISTFT) per utterance."""
    torch = H.torch
    from b200tts import capi, config, synth
    cfg, vcfg = config.F5, config.BIGVGAN
    ins = [synth.f5_inputs(1000 + rank * U + i, args.audio_len, args.n_text) for i in range(U)]
    N = int(ins[0][2][0])
    L = args.audio_len
    ref_len = L // cfg.hop + 1
    G = N - ref_len
    ns = cfg.hop * (G - 1)
    nv = vcfg.out_samples(G)
    work = f5_work(cfg, N, ref_len)
    # contiguous per-utterance blocks: audio [U][L] i16, text ids [U][n_text] i32, Euler-start noise [U][N*100] f32
    audio_h = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).pin_memory()
    ids_h = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).pin_memory()
    noise_h = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).pin_memory()
    audio_d, ids_d, noise_d = audio_h.cuda(), ids_h.cuda(), noise_h.cuda()
    pcm_d = torch.empty((U, nv if with_vocoder else ns), dtype=torch.int16, device="cuda")
    audio_np, ids_np, noise_np = audio_h.numpy(), ids_h.numpy(), noise_h.numpy().reshape(U, N, cfg.n_mels)
    wav_np = torch.empty((U, nv), dtype=torch.int16).pin_memory().numpy()
    torch.cuda.synchronize()

    def core():              # inputs resident in HBM, device-pointer entry point, no synchronisation
        if with_vocoder:
            eng.f5_bigvgan_pipeline_device(U, audio_d.data_ptr(), L, ids_d.data_ptr(), args.n_text, N, noise_d.data_ptr(),
                                           pcm_d.data_ptr(), precision=prec)
        else:
            eng.f5_synthesize_batch_device(U, audio_d.data_ptr(), L, ids_d.data_ptr(), args.n_text, N, noise_d.data_ptr(),
                                           pcm_d.data_ptr(), precision=prec)

    def step_e2e():          # the reference-facing host-buffer C call: H2D of audio / ids / noise and D2H of the PCM inside it
        if with_vocoder:
            eng.f5_bigvgan_pipeline(audio_np, ids_np, N, noise_np, precision=prec, out=wav_np)
        else:
            for u in range(U):
                eng.f5_synthesize(audio_np[u], ids_np[u], N, noise_np[u], precision=prec)

    with torch.cuda.stream(H.stream):
        for _ in range(warmup):
            core()
        step_e2e()
        step_e2e()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    ms = H.timed(core, steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_e2e = H.timed(step_e2e, steps)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(H.stream):
        core()
    host_enqueue_ms = 1e3 * (time.perf_counter() - t0)      # CPU time to enqueue one step (no sync): launch-bound check
    torch.cuda.synchronize()
    prof = profiled(eng, torch, H.stream, core)
    pk = peaks()
    frames = U * G * H.world * steps
    audio_s = U * (nv / vcfg.sample_rate if with_vocoder else work["audio_s"])
    h2d = U * (L * 2 + args.n_text * 4 + N * cfg.n_mels * 4)
    d2h = U * (nv * 2 if with_vocoder else ns * 2)
    res = {
        "value": frames / (ms / 1e3), "ms_per_step": ms / steps, "rtf": (ms / 1e3 / steps) / audio_s,
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "mel-frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / steps, "rtf": (ms_e2e / 1e3 / steps) / audio_s,
                "api": "b200tts_f5_bigvgan_pipeline (host buffers)" if with_vocoder else "b200tts_f5_synthesize (host buffers)"},
        "gpu_launches": int(launches), "clocks": clocks, "host_enqueue_ms_per_step": host_enqueue_ms,
        "profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
        "workload": (f"F5-TTS NFE={cfg.nfe} ({cfg.nfe - 1} Euler steps, CFG pair), {L / cfg.sample_rate:.0f} s ref / {args.n_text} text ids, "
                     f"N={N}, G={G} generated frames, {U} utterance(s) per GPU per step: preprocess + DiT loop + "
                     + ("BigVGAN-v2 24 kHz on the generated mel [BASELINE.json metric configuration: configs[3] per-GPU share]" if with_vocoder
                        else "Vocos/ISTFT decode [BASELINE.json configs[2]]")),
    }
    total_ms = max(sum(v["ms"] for v in prof.values()), 1e-9)
    kname = {capi.BF16: "bf16", capi.F16: "fp16"}.get(prec, "f32")
    chain = prof.get("f5.chain")
    if chain and chain["ms"] > 0:
        ach = U * work["steps"] * work["flops_chain_step"] / (chain["ms"] / 1e3) / 1e12
        res["roofline"] = {"bound": "tensor", "kernel": f"dit_chain_kernel ({kname}: out-proj, ff1, ff2 and the next q|k|v per launch, both LayerNorms folded into the GEMM epilogues)",
                           "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                           "traffic": ncu_traffic("f5.chain.team8") if U == 1 else None, "peak_source": pk["source"] + " (sustained cuBLAS bf16; fp16 has the same tensor rate)",
                           "avg_launch_ms": chain["ms"] / max(chain["launches"], 1), "share_of_step": chain["ms"] / total_ms}
    else:
        gemm_tags = ("f5.qkv_gemm", "f5.out_gemm", "f5.ff1_gemm", "f5.ff2_gemm")
        gemm_ms = sum(prof[t]["ms"] for t in gemm_tags if t in prof)
        gemm_n = sum(prof[t]["launches"] for t in gemm_tags if t in prof)
        if gemm_ms > 0:
            ach = U * work["steps"] * work["flops_gemm_step"] / (gemm_ms / 1e3) / 1e12
            res["roofline"] = {"bound": "tensor", "kernel": f"rowgemm_tc kernels ({kname}: DiT qkv/out/ff1/ff2 GEMMs)", "achieved": ach,
                               "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                               "traffic": ncu_traffic("f5.dit_gemm"), "peak_source": pk["source"] + " (sustained cuBLAS bf16)",
                               "avg_launch_ms": gemm_ms / max(gemm_n, 1), "share_of_step": gemm_ms / total_ms}
    att = prof.get("f5.attention")
    if att and att["ms"] > 0:
        ach = U * work["steps"] * work["flops_attn_step"] / (att["ms"] / 1e3) / 1e12
        res["roofline_attention"] = {"bound": "tensor", "kernel": "attn_tc_kernel", "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                                     "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                                     "avg_launch_ms": att["ms"] / max(att["launches"], 1), "share_of_step": att["ms"] / total_ms}
    res["dit_tflops_overall"] = U * work["flops_total"] / (ms / steps / 1e3) / 1e12
    res["dit_frac_of_tensor_peak"] = res["dit_tflops_overall"] / pk["bf16_tflops_sustained"]
    return res


def config4_set(cfg, n_utt=64, n_text=150):
    """BASELINE.json configs[3] / SURVEY.md 8d "Config 4": 64 utterances = config 3 with seeds 1000+i and reference lengths
    uniform in [4 s, 8 s] (N = 2 x reference frames in [752, 1502]). Returns [(seed, L)]."""
    rng = np.random.default_rng(4)
    secs = rng.uniform(4.0, 8.0, size=n_utt)
    return [(1000 + i, int(round(secs[i] * cfg.sample_rate))) for i in range(n_utt)]


def bench_config4(args, H, eng, rank, prec, steps, warmup, sampler=None):
    """STRONG scaling: the 64-utterance set is dealt to the ranks (longest first to the lightest rank, cost N^2 + c N); a rank
    runs its share as ONE ragged batch through b200tts_f5_bigvgan_pipeline_ragged (graph A per utterance, one DiT loop over all
    its 2 * sum(N) rows, BigVGAN per utterance). value = generated frames of ALL utterances / max-over-ranks time."""
    torch = H.torch
    from b200tts import config, distributed, synth
    cfg, vcfg = config.F5, config.BIGVGAN
    utts = config4_set(cfg, args.config4_utterances, args.n_text)
    Ns_all = [2 * (L // cfg.hop + 1) for _, L in utts]
    shards = distributed.shard_utterances([float(n) * n + 4096.0 * n for n in Ns_all], H.world)
    mine = shards[rank]
    ins = [synth.f5_inputs(utts[i][0], utts[i][1], args.n_text) for i in mine]
    L = np.asarray([utts[i][1] for i in mine], dtype=np.int64)
    Ns = np.asarray([Ns_all[i] for i in mine], dtype=np.int64)
    nt = np.full(len(mine), args.n_text, dtype=np.int32)
    G_mine = Ns - (L // cfg.hop + 1)
    G_all = [n - n // 2 for n in Ns_all]
    nv = int(sum(vcfg.out_samples(int(g)) for g in G_mine))
    audio_h = torch.from_numpy(np.concatenate([a.reshape(-1) for a, _, _, _ in ins])).pin_memory()
    ids_h = torch.from_numpy(np.concatenate([t.reshape(-1) for _, t, _, _ in ins])).pin_memory()
    noise_h = torch.from_numpy(np.concatenate([n.reshape(-1, cfg.n_mels) for _, _, _, n in ins], 0)).pin_memory()
    wav_h = torch.empty((nv,), dtype=torch.int16).pin_memory()
    audio_d, ids_d, noise_d = audio_h.cuda(), ids_h.cuda(), noise_h.cuda()
    wav_d = torch.empty((nv,), dtype=torch.int16, device="cuda")
    audio_np, ids_np, noise_np, wav_np = audio_h.numpy(), ids_h.numpy(), noise_h.numpy(), wav_h.numpy()
    U = len(mine)
    torch.cuda.synchronize()

    def core():
        eng.f5_bigvgan_pipeline_ragged_device(U, audio_d.data_ptr(), L, ids_d.data_ptr(), nt, Ns, noise_d.data_ptr(), wav_d.data_ptr(), precision=prec)

    def step_e2e():
        eng.f5_bigvgan_pipeline_ragged_concat(audio_np, L, ids_np, nt, Ns, noise_np, wav_np, precision=prec)

    with torch.cuda.stream(H.stream):
        for _ in range(warmup):
            core()
        step_e2e()
        step_e2e()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    ms = H.timed(core, steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_e2e = H.timed(step_e2e, steps)
    # per-rank busy time (one more step, timed locally, no barrier inside): the tail imbalance of the deal
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(H.stream):
        ev0.record(H.stream)
        core()
        ev1.record(H.stream)
    torch.cuda.synchronize()
    my_ms = ev0.elapsed_time(ev1)
    busy = [my_ms]
    if H.world > 1:
        t = torch.zeros(H.world, device="cuda")
        t[rank] = my_ms
        H.dist.all_reduce(t)
        busy = [float(v) for v in t.tolist()]
    prof = profiled(eng, torch, H.stream, core)
    pk = peaks()
    frames = float(sum(G_all)) * steps
    audio_s = sum(vcfg.out_samples(int(g)) for g in G_all) / vcfg.sample_rate
    flops_all = sum(f5_work(cfg, n, n // 2)["flops_total"] for n in Ns_all)
    h2d = int(audio_h.numel() * 2 + ids_h.numel() * 4 + noise_h.numel() * 4)
    res = {
        "value": frames / (ms / 1e3), "ms_per_step": ms / steps, "rtf": (ms / 1e3 / steps) / audio_s,
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "mel-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(nv * 2),
                "ms_per_step": ms_e2e / steps, "rtf": (ms_e2e / 1e3 / steps) / audio_s, "api": "b200tts_f5_bigvgan_pipeline_ragged (host buffers)",
                "bytes_are": "this rank's share"},
        "gpu_launches": int(launches), "clocks": clocks,
        "profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
        "rank_busy_ms": [round(b, 2) for b in busy], "tail_imbalance": (max(busy) / (sum(busy) / len(busy)) - 1.0) if busy else None,
        "utterances_per_rank": [len(s) for s in shards],
        "workload": (f"F5-TTS NFE={cfg.nfe} + BigVGAN-v2 24 kHz, batch of {len(utts)} synthetic utterances (seeds 1000+i, references uniform in "
                     f"[4 s, 8 s] -> N in [{min(Ns_all)}, {max(Ns_all)}], {args.n_text} text ids) dealt longest-first to {H.world} GPU(s); each "
                     "rank runs its share as ONE ragged batch [BASELINE.json configs[3]]"),
    }
    total_ms = max(sum(v["ms"] for v in prof.values()), 1e-9)
    kname = {2: "fp16", 1: "bf16"}.get(int(prec), "f32")
    chain = prof.get("f5.chain")
    if chain and chain["ms"] > 0:
        steps_n = cfg.nfe - 1
        fl = sum(f5_work(cfg, int(n), int(n) // 2)["flops_chain_step"] for n in Ns) * steps_n
        ach = fl / (chain["ms"] / 1e3) / 1e12
        res["roofline"] = {"bound": "tensor", "kernel": f"dit_chain_kernel ({kname}: out-proj, ff1, ff2 and the next q|k|v per launch, both LayerNorms folded into the GEMM epilogues)",
                           "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                           "traffic": ncu_traffic("f5.chain"), "peak_source": pk["source"] + " (sustained cuBLAS bf16; fp16 has the same tensor rate)",
                           "avg_launch_ms": chain["ms"] / max(chain["launches"], 1), "share_of_step": chain["ms"] / total_ms, "rank": 0}
    res["dit_tflops_overall"] = flops_all / (ms / steps / 1e3) / 1e12
    res["dit_frac_of_tensor_peak"] = res["dit_tflops_overall"] / (pk["bf16_tflops_sustained"] * H.world)
    return res


def bench_indextts_vocoder(args, H, eng, rank, prec, steps, warmup):
    """Vocoder half of BASELINE.json configs[4] (IndexTTS_F, S latent rows -> 1024*(S-2)+30 samples) through the host-buffer
    C ABI call a reference script would make: H2D of the latent and conditioning vectors, D2H of the PCM inside the timed
    region. There is no device-pointer entry point for this graph, so `value` and `e2e` are the same measurement."""
    torch = H.torch
    from b200tts import config, synth
    cfg = config.INDEXTTS_VOCODER
    S = args.latent_rows
    conds, cond_layer, hidden = synth.ivgan_inputs(300 + rank, S)

    def step():
        eng.indextts_vocoder_run(hidden, conds, cond_layer, precision=prec, hop=cfg.hop)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    H.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()                                       # synchronises before returning (host-pointer entry point)
    dt = time.perf_counter() - t0
    ms = 1e3 * dt
    if H.world > 1:
        t = torch.tensor([ms], device="cuda")
        H.dist.all_reduce(t, op=H.dist.ReduceOp.MAX)
        ms = float(t.item())
    frames = (S - 2) * 4 * H.world * steps              # one latent row = 1024 samples = 4 frames of 256 samples
    n_out = cfg.out_samples(S - 2)
    e2e = {"value": frames / (ms / 1e3), "unit": "mel-frames/s", "ms_per_step": ms / steps,
           "h2d_bytes_per_step": int(hidden[:-2].nbytes + sum(c.nbytes for c in conds) + cond_layer.nbytes),
           "d2h_bytes_per_step": int(n_out * 2), "rtf": (ms / 1e3 / steps) / (n_out / cfg.sample_rate)}
    return {"value": e2e["value"], "ms_per_step": ms / steps, "rtf": e2e["rtf"], "e2e": e2e, "gpu_launches": None, "clocks": None,
            "profile_ms": {},
            "workload": f"IndexTTS_F vocoder (BigVGAN x1024 with conditioning), latent (S={S}, 1280) -> int16 PCM (1,1,{n_out}), one "
                        "utterance per step through the host-buffer call (timed on the host clock around synchronising calls) "
                        "[vocoder half of BASELINE.json configs[4]]"}


def bench_indextts_gpt(args, H, eng, rank, prec, steps, warmup, sampler=None):
    """Acoustic half of BASELINE.json configs[4]: one sentence of IndexTTS GPT-2 greedy decode per step -- graphs B, C, D,
    a prefill of (32 conditioning latents + text + 3) rows and `--new-tokens` single-row decode calls with the KV cache,
    penalty window and loop state on the device. Metric: generated mel tokens per second."""
    torch = H.torch
    from b200tts import capi, config, synth
    cfg = config.INDEXTTS_GPT
    n_new = args.new_tokens
    conds, text_ids = synth.igpt_inputs(900 + rank, args.gpt_text, cfg)
    D, cap = cfg.dim, cfg.max_generate + 1
    rows = cfg.cond_rows + args.gpt_text + 3
    conds_h = torch.from_numpy(conds.reshape(-1, D)).pin_memory()
    ids_h = torch.from_numpy(text_ids.reshape(-1)).pin_memory()
    conds_d, ids_d = conds_h.cuda(), ids_h.cuda()
    out_ids = torch.zeros((cap,), dtype=torch.int32, device="cuda")
    out_hid = torch.zeros((cap, D), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    produced = []

    def core():
        produced.append(eng.indextts_gpt_generate_device(conds_d.data_ptr(), cfg.cond_rows, ids_d.data_ptr(), args.gpt_text,
                                                         out_ids.data_ptr(), out_hid.data_ptr(), max_new=n_new, precision=prec))

    def step_e2e():
        eng.indextts_gpt_generate(conds, text_ids, max_new=n_new, precision=prec)       # host buffers in and out, synchronises

    with torch.cuda.stream(H.stream):
        for _ in range(warmup):
            core()
        step_e2e()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    produced.clear()
    ms = H.timed(core, steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    tokens = sum(produced)
    ms_e2e = H.timed(step_e2e, steps)
    # per-kernel split of ONE sentence (eager, CUDA events around every launch)
    eng.profile_begin()
    with torch.cuda.stream(H.stream):
        core()
    prof = eng.profile_end()
    pk = peaks()
    wbytes = 2 if prec == capi.BF16 else 4
    per_tok_w = (cfg.layers * 12 * D * D + cfg.mel_codes * D) * wbytes                      # every projection weight once
    avg_kv = rows + n_new / 2
    per_tok_kv = cfg.layers * 2 * avg_kv * D * 4                                            # fp32 cache rows read by attention
    gemv_tags = ("igpt.qkv_gemv", "igpt.out_gemv", "igpt.fc_gemv", "igpt.proj_gemv", "igpt.head")
    gemv_ms = sum(prof[t]["ms"] for t in gemv_tags if t in prof)
    gemv_n = sum(prof[t]["launches"] for t in gemv_tags if t in prof)
    total_ms = max(sum(v["ms"] for v in prof.values()), 1e-9)
    n_tok = produced[-1]
    res = {
        "value": tokens * H.world / (ms / 1e3), "ms_per_step": ms / steps,
        "e2e": {"value": tokens * H.world / (ms_e2e / 1e3), "unit": "tokens/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(conds.nbytes + text_ids.nbytes), "d2h_bytes_per_step": int(n_tok * (D * 4 + 4))},
        "gpu_launches": int(launches), "clocks": clocks, "profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
        "ms_per_token": ms / max(tokens, 1),
        "workload": (f"IndexTTS GPT-2 acoustic model ({cfg.layers} layers, dim {D}, {cfg.heads} heads, {cfg.mel_codes} mel codes), one sentence "
                     f"per step: prefill of {rows} rows ({cfg.cond_rows} latents + {args.gpt_text} text ids + 3) then {n_new - 1} single-row "
                     f"greedy decode calls, repeat-penalty window on the device [acoustic half of BASELINE.json configs[4]]"),
    }
    pers = prof.get("igpt.decode_persistent")
    if pers and pers["ms"] > 0:
        # the persistent decode kernel: every decode call streams every projection weight once and reads the cache rows of its head
        ach = (n_tok - 1) * (per_tok_w + per_tok_kv) / (pers["ms"] / 1e3) / 1e9
        res["roofline"] = {"bound": "hbm", "kernel": "gpt_decode_kernel (persistent: 24 layers + head + pick per token, up to 32 tokens per launch)",
                           "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": ncu_traffic("igpt.decode"),
                           "peak_source": pk["source"], "avg_launch_ms": pers["ms"] / max(pers["launches"], 1),
                           "share_of_step": pers["ms"] / total_ms, "bytes_per_token": int(per_tok_w), "kv_bytes_per_token_avg": int(per_tok_kv),
                           "ms_per_token_in_kernel": pers["ms"] / max(n_tok - 1, 1),
                           "traffic_note": "ncu capture of a 7-token launch (profiles/r01/traffic.json): 1.0 GB read per token"}
    elif gemv_ms > 0:
        # per-kernel decode path (B200TTS_GPT_PERSIST=0): (n_tok - 1) decode calls stream every weight once; the prefill's head call too
        ach = ((n_tok - 1) * per_tok_w + cfg.mel_codes * D * wbytes) / (gemv_ms / 1e3) / 1e9
        res["roofline"] = {"bound": "hbm", "kernel": "gemv_kernel (decode projections: LayerNorm + matrix-vector + epilogue)", "achieved": ach,
                           "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": ncu_traffic("igpt.gemv"),
                           "peak_source": pk["source"], "avg_launch_ms": gemv_ms / max(gemv_n, 1), "share_of_step": gemv_ms / total_ms,
                           "bytes_per_token": int(per_tok_w), "kv_bytes_per_token_avg": int(per_tok_kv)}
    res["hbm_floor_ms_per_token"] = (per_tok_w + per_tok_kv) / (pk["hbm_gbs"] * 1e9) * 1e3
    return res


def bench_indextts(args, H, eng, rank, prec, steps, warmup, sampler=None):
    """BASELINE.json configs[4] end to end for one sentence per step, through the host-buffer calls a reference script would make
    (Inference_IndexTTS_ONNX.py:719-791 minus the conditioning graph A, whose outputs are inputs here): GPT-2 greedy decode of
    `--new-tokens` mel tokens (graphs B, C, D, E) -> IndexTTS_F vocoder on the hidden states -> int16 PCM. Every call takes host
    pointers and synchronises, so `value` and `e2e` are the same measurement (timed on the host clock)."""
    torch = H.torch
    from b200tts import config, synth
    gcfg, vcfg = config.INDEXTTS_GPT, config.INDEXTTS_VOCODER
    conds, text_ids = synth.igpt_inputs(900 + rank, args.gpt_text, gcfg)
    vconds, cond_layer, _ = synth.ivgan_inputs(300 + rank, 3)
    n_new = args.new_tokens
    out = {}

    def step():
        ids, hidden = eng.indextts_gpt_generate(conds, text_ids, max_new=n_new, precision=prec)
        out["n"] = len(ids)
        out["pcm"] = eng.indextts_vocoder_run(hidden, vconds, cond_layer, precision=prec, hop=vcfg.hop)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    H.barrier()
    l0 = eng.launch_count()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count() - l0
    ms = 1e3 * dt
    if H.world > 1:
        t = torch.tensor([ms], device="cuda")
        H.dist.all_reduce(t, op=H.dist.ReduceOp.MAX)
        ms = float(t.item())
    n_tok = out["n"]
    n_samp = int(out["pcm"].shape[-1])
    frames = (n_tok - 2) * 4 * H.world * steps              # one latent row = 1024 samples = 4 frames of 256 samples
    audio_s = n_samp / vcfg.sample_rate
    e2e = {"value": frames / (ms / 1e3), "unit": "mel-frames/s", "ms_per_step": ms / steps, "rtf": (ms / 1e3 / steps) / audio_s,
           "h2d_bytes_per_step": int(conds.nbytes + text_ids.nbytes + n_tok * gcfg.dim * 4 + sum(c.nbytes for c in vconds) + cond_layer.nbytes),
           "d2h_bytes_per_step": int(n_tok * (gcfg.dim * 4 + 4) + n_samp * 2)}
    return {"value": e2e["value"], "ms_per_step": ms / steps, "rtf": e2e["rtf"], "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "profile_ms": {}, "tokens_per_s": n_tok * H.world * steps / (ms / 1e3),
            "workload": f"IndexTTS sentence end to end without the conditioning graph: GPT-2 decode of {n_tok} mel tokens (prompt "
                        f"{gcfg.cond_rows + args.gpt_text + 3} rows) + IndexTTS_F vocoder -> {n_samp} samples ({audio_s:.2f} s) "
                        "[BASELINE.json configs[4]]; host-buffer calls, host clock"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4", choices=["config4", "pipeline", "f5", "bigvgan", "indextts_vocoder", "indextts_gpt", "indextts"])
    ap.add_argument("--config4-utterances", type=int, default=64)
    ap.add_argument("--new-tokens", type=int, default=256, help="indextts_gpt workload: E calls per sentence (prefill + decode)")
    ap.add_argument("--gpt-text", type=int, default=60, help="indextts_gpt workload: text ids per sentence")
    ap.add_argument("--latent-rows", type=int, default=142, help="indextts_vocoder workload: rows of save_hidden_state")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--audio-len", type=int, default=144000)
    ap.add_argument("--n-text", type=int, default=150)
    ap.add_argument("--utterances", type=int, default=8, help="pipeline workload: utterances per GPU per step")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="operand type of the tensor-core engine (BASELINE.json's configurations name fp16); fp32 = the SIMT parity engine")
    ap.add_argument("--no-chain", action="store_true", help="DiT blocks as separate launches instead of the fused row-block chain")
    ap.add_argument("--fp8", type=int, nargs="?", const=2, default=0, choices=[0, 1, 2],
                    help="optional lower-fidelity mode of the fused chain (NOT the default): 1 = ff1 and q|k|v with e4m3 operands (PCM SNR "
                         "~32 dB against the fp32 reference instead of ~62 dB), 2 (bare --fp8) = ff2 as well (~30.5 dB)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the attached f5 / bigvgan measurements of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import b200tts  # noqa: F401
    from b200tts import capi, config, distributed, synth, weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libb200tts has no CPU fallback")
    torch.cuda.set_device(local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL writes its banner ("NCCL version ...") to fd 1 when the communicator is created; stdout must carry the ONE JSON
        # line only, so fd 1 points at stderr until the first collectives (the weight broadcast) are done.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        t = torch.zeros(1, device="cuda")
        dist.all_reduce(t)
        torch.cuda.synchronize()

    prec = {"fp16": capi.F16, "bf16": capi.BF16, "fp32": capi.F32}[args.precision]
    if args.workload in ("indextts_gpt", "indextts", "indextts_vocoder") and prec == capi.F16:
        prec = capi.BF16                                       # the IndexTTS entry points take bf16 weights
    dtype = {capi.F16: "fp16", capi.BF16: "bf16", capi.F32: "f32"}[prec]
    eng = capi.Engine(local_rank)
    if args.no_chain:
        eng.set_option("dit_chain", 0)
    if args.fp8:
        eng.set_option("dit_fp8", args.fp8)
        dtype += "+e4m3(ff1,qkv)" if args.fp8 == 1 else "+e4m3(ff1,ff2,qkv)"
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    H = Harness(torch, dist, stream, world)
    need_f5 = args.workload in ("f5", "pipeline", "config4")
    need_vgan = args.workload in ("bigvgan", "pipeline", "config4")
    if args.workload in ("indextts_vocoder", "indextts"):
        cfgv = config.INDEXTTS_VOCODER
        state = weights.ivgan_engine_tensors(synth.ivgan_state(777), cfgv) if rank == 0 else None
        distributed.load_state_broadcast(eng, "ivgan", state, src=0)
        eng.indextts_vocoder_build()

    if args.workload in ("indextts_gpt", "indextts"):
        state = weights.igpt_engine_tensors(synth.igpt_state(555), config.INDEXTTS_GPT) if rank == 0 else None
        distributed.load_state_broadcast(eng, "igpt", state, src=0)
        eng.indextts_gpt_build()

    # weights: rank 0 makes them, NCCL broadcast over NVLink to the others (the only collective of the job)
    if need_vgan:
        state = weights.bigvgan_engine_tensors(synth.bigvgan_state(1234)) if rank == 0 else None
        distributed.load_state_broadcast(eng, "bigvgan", state, src=0)
        eng.bigvgan_build()
    if need_f5:
        cfg5 = config.F5
        if rank == 0:
            dsd = synth.f5_dit_state(4321)
            parts = {"dit": weights.dit_engine_tensors(dsd, cfg5), "vocos": weights.vocos_engine_tensors(synth.vocos_state(2468), cfg5),
                     "f5": weights.f5_export_constants(dsd, cfg5)}
        else:
            parts = {"dit": None, "vocos": None, "f5": None}
        for k in ("dit", "vocos", "f5"):
            distributed.load_state_broadcast(eng, k, parts[k], src=0)
        eng.f5_build()

    if saved_stdout is not None:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    extra = {}
    if args.workload == "bigvgan":
        res = bench_bigvgan(args, H, eng, rank, args.batch, args.frames, prec, args.steps, args.warmup, sampler)
    elif args.workload == "indextts_vocoder":
        res = bench_indextts_vocoder(args, H, eng, rank, prec, args.steps, args.warmup)
    elif args.workload == "indextts":
        res = bench_indextts(args, H, eng, rank, prec, args.steps, args.warmup, sampler)
    elif args.workload == "indextts_gpt":
        res = bench_indextts_gpt(args, H, eng, rank, prec, args.steps, args.warmup, sampler)
        dtype = "bf16 weights, fp32 activations / cache / accumulation" if prec == capi.BF16 else "f32"
    elif args.workload == "f5":
        res = bench_f5(args, H, eng, rank, prec, args.steps, args.warmup, sampler=sampler)
    elif args.workload == "config4":
        res = bench_config4(args, H, eng, rank, prec, args.steps, args.warmup, sampler=sampler)
        if not args.no_extras and world == 1:
            # the other single-GPU configurations, short: the uniform pipeline (8 x config-3 utterances per GPU), configs[2] (one
            # utterance, latency) and configs[1] (vocoder alone)
            ppr = bench_f5(args, H, eng, rank, prec, steps=3, warmup=3, with_vocoder=True, U=args.utterances)
            extra["pipeline"] = {"metric": "mel_frames_per_s", "unit": "mel-frames/s", **ppr}
            f5r = bench_f5(args, H, eng, rank, prec, steps=5, warmup=3)
            extra["f5"] = {"metric": "mel_frames_per_s", "unit": "mel-frames/s", **f5r}
            vgr = bench_bigvgan(args, H, eng, rank, args.batch, args.frames, prec, steps=10, warmup=3)
            extra["bigvgan"] = {"metric": "mel_frames_per_s", "unit": "mel-frames/s", **vgr}
    else:
        res = bench_f5(args, H, eng, rank, prec, args.steps, args.warmup, with_vocoder=True, U=args.utterances, sampler=sampler)

    if rank == 0:
        workload = res.pop("workload")
        line = {"metric": "mel_frames_per_s", "value": res.pop("value"), "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "strong" if args.workload == "config4" else "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": workload, "parallelism": f"dp{world} (utterance sharding, weights NCCL-broadcast at load)",
                           "l2": "per-step working set (activations + weights, > 0.5 GB) exceeds the 126 MB L2; no explicit flush"}}
        if args.workload == "indextts_gpt":
            line["metric"], line["unit"] = "mel_tokens_per_s", "tokens/s"
        line.update(res)
        line.update(extra)
        if not args.no_cpu_baseline and world == 1:              # rank 0 at N = 1 only: the other ranks would idle behind it
            cores = os.cpu_count() or 1
            if args.workload == "indextts_vocoder":
                import torch as _t
                from oracle import indextts_ref
                _t.set_num_threads(cores)
                sdv = synth.ivgan_state(777)
                cds, cl, hid = synth.ivgan_inputs(300, args.latent_rows)
                indextts_ref.indextts_f_pcm(hid[:6], cds, cl, sdv, config.INDEXTTS_VOCODER)
                t0 = time.perf_counter()
                indextts_ref.indextts_f_pcm(hid, cds, cl, sdv, config.INDEXTTS_VOCODER)
                dtc = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": (args.latent_rows - 2) * 4 / dtc, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                                        "sample": f"1 x latent ({args.latent_rows},1280), oracle (torch-CPU fp32 restatement of IndexTTS_F)",
                                        "s_per_utterance": dtc}
            elif args.workload == "indextts":
                pass                                           # see --workload indextts_gpt / indextts_vocoder for the two CPU legs
            elif args.workload == "indextts_gpt":
                import torch as _t
                from oracle import indextts_gpt_ref
                _t.set_num_threads(cores)
                cfgg = config.INDEXTTS_GPT
                sdg = synth.igpt_state(555)
                cds, tid = synth.igpt_inputs(900, args.gpt_text, cfgg)
                t0 = time.perf_counter()
                indextts_gpt_ref.generate(cds, tid, sdg, cfgg, max_new=1)
                t_pre = time.perf_counter() - t0
                n_cpu = 9
                t0 = time.perf_counter()
                indextts_gpt_ref.generate(cds, tid, sdg, cfgg, max_new=n_cpu)
                dtc = time.perf_counter() - t0
                per_tok = max(dtc - t_pre, 1e-9) / (n_cpu - 1)
                line["cpu_baseline"] = {"value": args.new_tokens / (t_pre + (args.new_tokens - 1) * per_tok), "unit": "tokens/s", "cores": cores,
                                        "kind": "port", "sample": f"prefill + {n_cpu - 1} decode calls of one sentence (extrapolated to "
                                        f"{args.new_tokens}); oracle (torch-CPU fp32 restatement of graphs B-E)",
                                        "prefill_s": t_pre, "s_per_decode_call": per_tok}
            elif args.workload == "bigvgan":
                dt = cpu_bigvgan(args.frames, 2, cores)
                line["cpu_baseline"] = {"value": args.frames / dt, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                                        "sample": f"2 x 1 mel (1,100,{args.frames}), oracle (torch-CPU fp32 restatement of the reference "
                                                  "modules; ORT is not installable offline)", "s_per_mel": dt}
            else:
                r = cpu_f5(args.audio_len, args.n_text, 2, cores)
                G = r["N"] - r["ref_len"]
                tot = r["pre_s"] + (config.F5.nfe - 1) * r["step_s"]
                if args.workload in ("pipeline", "config4"):
                    r["bigvgan_s"] = cpu_bigvgan(G, 1, cores)
                    tot += r["bigvgan_s"]
                else:
                    tot += r["decode_s"]
                line["cpu_baseline"] = {"value": G / tot, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                                        "sample": f"ONE utterance of the batch: graph A, 2 of {config.F5.nfe - 1} DiT steps (extrapolated), "
                                                  + ("BigVGAN on its generated frames" if args.workload in ("pipeline", "config4") else "graph C")
                                                  + "; oracle (torch-CPU fp32 restatement of the reference modules; ORT is not installable offline)",
                                        "parts_s": r}
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the B200 TTS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload bigvgan|f5]

One "step" = one pass of the hot path over one batch of synthetic input. Default workload (N = 1): BASELINE.json
configs[1], BigVGAN-v2 24khz_100band_256x on mels (8, 100, 512), tensor-core (bf16 operand / fp32 accumulate) path.
Each rank runs the same per-GPU batch (utterances shard with no data-path collective: weak scaling); NCCL is used
once, to broadcast the weights from rank 0 at load. Prints ONE JSON line on rank 0.
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

# algorithmic work per mel frame (DESIGN.md "Work model")
BIGVGAN_RESCONV_MFLOP_PER_FRAME = 1746.47      # the 108 dilated resblock convs
BIGVGAN_TOTAL_MFLOP_PER_FRAME = 1804.0         # + conv_pre, upsamplers, conv_post
BIGVGAN_AA_ELEMS_PER_FRAME = 18 * 2 * 33792 / 6.0 * 0 + 0  # placeholder, computed below


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


def bigvgan_work(cfg, B, T):
    """Algorithmic FLOPs / bytes of one step (B mels of T frames)."""
    frames = B * T
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
    return {"frames": frames, "flops_resconv": B * flops_res, "flops_total": B * (flops_res + flops_up + flops_pre + flops_post),
            "aa_elems": B * aa_elems, "audio_s": B * cfg.out_samples(T) / cfg.sample_rate}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (the oracle restatement of its PyTorch modules -- onnxruntime is
    not installable offline) on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    import torch
    import b200tts  # noqa: F401
    from b200tts import config, synth
    from oracle import bigvgan_ref
    cfg = config.BIGVGAN
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, T = 1, args.frames                      # sample: ONE mel of the (8,100,512) batch per step
    sd = synth.bigvgan_state(1234)
    mel = synth.bigvgan_mel(100, B, T)
    for _ in range(args.warmup):
        bigvgan_ref.bigvgan_pcm(mel, sd, cfg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bigvgan_ref.bigvgan_pcm(mel, sd, cfg)
    dt = time.perf_counter() - t0
    v = B * T * args.steps / dt
    line = {
        "impl": "reference", "metric": "mel_frames_per_s", "value": v, "unit": "mel-frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BigVGAN-v2 24khz_100band_256x, mels ({args.batch},100,{T}) [configs[1]]; "
                               f"each reference step = 1 mel of that batch", "parallelism": "cpu"},
        "cpu_baseline": {"value": v, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                         "sample": f"1 mel (1,100,{T}) per step, torch-CPU fp32 eager restatement of the reference modules "
                                   f"(stand-in for ORT CPUExecutionProvider)"},
        "e2e": {"value": v, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rtf": dt / args.steps / (B * cfg.out_samples(T) / cfg.sample_rate),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bigvgan", choices=["bigvgan"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5                     # ~6 s of CPU work per step: keep the arm within minutes
        args.warmup = min(args.warmup, 1)
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import b200tts  # noqa: F401
    from b200tts import capi, config, distributed, synth, weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libb200tts has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = config.BIGVGAN
    prec = capi.BF16 if args.precision == "bf16" else capi.F32
    eng = capi.Engine(local_rank)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)

    # weights: rank 0 makes them, NCCL broadcast over NVLink to the others (the only collective of the job)
    state = weights.bigvgan_engine_tensors(synth.bigvgan_state(1234)) if rank == 0 else None
    distributed.load_state_broadcast(eng, "bigvgan", state, src=0)
    eng.bigvgan_build()

    B, T = args.batch, args.frames
    n_out = cfg.out_samples(T)
    work = bigvgan_work(cfg, B, T)
    mel_host = torch.from_numpy(synth.bigvgan_mel(100 + rank, B, T)).pin_memory()
    pcm_host = torch.empty((B, 1, n_out), dtype=torch.int16).pin_memory()
    mel_dev = mel_host.cuda(non_blocking=True)
    pcm_dev = torch.empty((B, 1, n_out), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()

    def step_device():
        eng.bigvgan_run_device(mel_dev.data_ptr(), B, T, pcm_dev.data_ptr(), precision=prec)

    def step_e2e():
        with torch.cuda.stream(stream):
            mel_dev.copy_(mel_host, non_blocking=True)
            eng.bigvgan_run_device(mel_dev.data_ptr(), B, T, pcm_dev.data_ptr(), precision=prec)
            pcm_host.copy_(pcm_dev, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(steps):
                fn()
            ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
            step_e2e()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ms = timed(step_device, args.steps)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps)

    # per-kernel CUDA-event timing of one more step (roofline leg; not part of the throughput number)
    eng.profile_begin()
    with torch.cuda.stream(stream):
        step_device()
    prof = eng.profile_end()

    if rank == 0:
        pk = peaks()
        frames_job = work["frames"] * world * args.steps
        value = frames_job / (ms / 1e3)
        e2e = frames_job / (ms_e2e / 1e3)
        conv = prof.get("bigvgan.resconv", {"launches": 0, "ms": 0.0})
        aa = prof.get("bigvgan.aa_snake", {"launches": 0, "ms": 0.0})
        line = {
            "metric": "mel_frames_per_s", "value": value, "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if prec == capi.BF16 else "f32", "data": "synthetic",
            "config": {"workload": f"BigVGAN-v2 24khz_100band_256x, mels ({B},100,{T}) per GPU -> int16 PCM ({B},1,{n_out}) "
                                   f"[BASELINE.json configs[1]]",
                       "parallelism": f"dp{world} (utterance sharding, weights NCCL-broadcast at load)",
                       "l2": "per-step working set ~0.6 GB >> 126 MB L2, no explicit flush"},
            "rtf": (ms / 1e3 / args.steps) / work["audio_s"],
            "e2e": {"value": e2e, "unit": "mel-frames/s", "h2d_bytes_per_step": int(mel_host.numel() * 4),
                    "d2h_bytes_per_step": int(pcm_host.numel() * 2), "ms_per_step": ms_e2e / args.steps,
                    "rtf": (ms_e2e / 1e3 / args.steps) / work["audio_s"]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
        }
        if conv["ms"] > 0:
            if prec == capi.BF16:
                ach = work["flops_resconv"] / (conv["ms"] / 1e3) / 1e12
                line["roofline"] = {"bound": "tensor", "kernel": "rowgemm_tc_kernel (108 resblock convs)",
                                    "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
                                    "peak_source": pk["source"] + " (sustained cuBLAS bf16)",
                                    "avg_launch_ms": conv["ms"] / max(conv["launches"], 1)}
            else:
                ach = work["flops_resconv"] / (conv["ms"] / 1e3) / 1e12
                line["roofline"] = {"bound": "tensor", "kernel": "rowgemm_f32_kernel (SIMT parity engine)", "achieved": ach,
                                    "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": pk["source"]}
        if aa["ms"] > 0:
            bytes_aa = work["aa_elems"] * (4 + 2 if prec == capi.BF16 else 8) * 0  # refined below
            # aa_snake reads fp32 or bf16 and writes bf16 (fast path): 108 launches, half read fp32 (4 B) and half bf16 (2 B)
            per_elem = (3.0 + 2.0) if prec == capi.BF16 else 8.0
            gbs = work["aa_elems"] * per_elem / (aa["ms"] / 1e3) / 1e9
            line["roofline_hbm"] = {"bound": "hbm", "kernel": "aa_snake_kernel", "achieved": gbs, "peak": pk["hbm_gbs"],
                                    "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None,
                                    "avg_launch_ms": aa["ms"] / max(aa["launches"], 1)}
        if not args.no_cpu_baseline:
            import torch as _t
            from oracle import bigvgan_ref
            cores = os.cpu_count() or 1
            _t.set_num_threads(cores)
            sd = synth.bigvgan_state(1234)
            m1 = synth.bigvgan_mel(100, 1, T)
            bigvgan_ref.bigvgan_pcm(m1[:, :, :32], sd, cfg)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                bigvgan_ref.bigvgan_pcm(m1, sd, cfg)
            dt = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": T / dt, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                                    "sample": f"{reps} x 1 mel (1,100,{T}), oracle (torch-CPU fp32 restatement of the reference "
                                              f"modules; ORT is not installable offline)", "s_per_mel": dt}
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

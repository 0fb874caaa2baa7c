"""Profiling driver (run under ncu on the GPU box): one F5 utterance of config-3 shape with a few Euler steps,
or one BigVGAN pass. Not a benchmark: numbers printed under a profiler are never bench values."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200tts  # noqa: F401,E402
from b200tts import capi, config, synth, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="f5", choices=["f5", "bigvgan", "igpt"])
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
eng = capi.Engine(0)
if args.what == "f5":
    cfg = config.F5
    dsd = synth.f5_dit_state(4321)
    eng.load_state("dit", weights.dit_engine_tensors(dsd, cfg))
    eng.load_state("vocos", weights.vocos_engine_tensors(synth.vocos_state(2468), cfg))
    eng.load_state("f5", weights.f5_export_constants(dsd, cfg))
    eng.f5_build()
    audio, ids, maxd, noise = synth.f5_inputs(1)
    for _ in range(args.reps + 1):      # first call builds the bf16 weight layouts; profile the later ones
        pcm = eng.f5_synthesize(audio, ids, int(maxd[0]), noise, precision=capi.BF16, n_steps=args.steps)
    print("f5 ok", pcm.shape, eng.launch_count())
elif args.what == "igpt":
    cfg = config.INDEXTTS_GPT
    eng.load_state("igpt", weights.igpt_engine_tensors(synth.igpt_state(555), cfg))
    eng.indextts_gpt_build()
    conds, ids = synth.igpt_inputs(900, 60, cfg)
    for _ in range(args.reps + 1):      # first call builds the bf16 weight layouts
        out = eng.indextts_gpt_generate(conds, ids, max_new=args.steps, precision=capi.BF16)
    print("igpt ok", out[0][:8], eng.launch_count())
else:
    eng.load_state("bigvgan", weights.bigvgan_engine_tensors(synth.bigvgan_state(1234)))
    eng.bigvgan_build()
    mel = synth.bigvgan_mel(100, args.batch, args.frames)
    for _ in range(args.reps + 1):
        pcm = eng.bigvgan_run(mel, precision=capi.BF16)
    print("bigvgan ok", pcm.shape, eng.launch_count())

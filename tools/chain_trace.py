"""Timeline of the fused DiT chain kernel (dit_chain.cu) from its %globaltimer stamps.

    B200TTS_GRAPHS=0 B200TTS_CHAIN_TRACE=gpurun_out/chain_trace.bin python tools/chain_trace.py [U]

Runs two Euler steps of a config-3 utterance (the stamps of the LAST chain launch survive) and prints, per stamp, the
min / median / max over the CTAs in microseconds from the first stamp of the kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402,F401
from b200tts import capi, config, synth, weights  # noqa: E402

NAMES = {0: "epi: kernel body starts", 1: "epi: pdl_wait passed"}
for j, jn in enumerate(["out", "ff1", "ff2", "qkv"]):
    b = 8 + 8 * j
    NAMES.update({b + 0: f"{jn}: A producer reaches the job", b + 1: f"{jn}: A hand-off passed, first TMA", b + 2: f"{jn}: MMA first chunk ready",
                  b + 3: f"{jn}: MMA last chunk issued", b + 4: f"{jn}: epilogue sees acc_full", b + 5: f"{jn}: epilogue tile done",
                  b + 6: f"{jn}: LN statistics of the team in", b + 7: f"{jn}: slab published"})


def main():
    U = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    path = os.environ.get("B200TTS_CHAIN_TRACE")
    import torch
    eng = capi.Engine(0)
    cfg = config.F5
    dsd = synth.f5_dit_state(4321)
    eng.load_state("dit", weights.dit_engine_tensors(dsd, cfg))
    eng.load_state("vocos", weights.vocos_engine_tensors(synth.vocos_state(2468), cfg))
    eng.load_state("f5", weights.f5_export_constants(dsd, cfg))
    eng.f5_build()
    L, n_text = 144000, 150
    ins = [synth.f5_inputs(1000 + i, L, n_text) for i in range(U)]
    N = int(ins[0][2][0])
    ns = 256 * (N - (L // 256 + 1) - 1)
    audio = torch.from_numpy(np.stack([a.reshape(-1) for a, _, _, _ in ins])).cuda()
    ids = torch.from_numpy(np.stack([t.reshape(-1) for _, t, _, _ in ins])).cuda()
    noise = torch.from_numpy(np.stack([n.reshape(-1) for _, _, _, n in ins])).cuda()
    pcm = torch.zeros((U, ns), dtype=torch.int16, device="cuda")
    prec = capi.F16 if os.environ.get("PREC", "f16") == "f16" else capi.BF16
    for _ in range(2):
        eng.f5_synthesize_batch_device(U, audio.data_ptr(), L, ids.data_ptr(), n_text, N, noise.data_ptr(), pcm.data_ptr(), precision=prec,
                                       n_steps=2)
        eng.synchronize()
    if not path:
        return
    t = np.fromfile(path, dtype=np.uint64).reshape(-1, 64).astype(np.float64)
    used = t[:, 0] > 0
    t = t[used]
    t0 = t[t > 0].min()
    print(f"U={U} CTAs traced: {t.shape[0]}")
    print(f"{'stamp':44s} {'n':>4s} {'min':>8s} {'median':>8s} {'max':>8s}   (us from the first stamp)")
    for slot in sorted(NAMES):
        v = t[:, slot]
        v = v[v > 0]
        if v.size == 0:
            continue
        v = (v - t0) / 1e3
        print(f"{NAMES[slot]:44s} {v.size:4d} {v.min():8.2f} {np.median(v):8.2f} {v.max():8.2f}")
    # one team in detail: CTA 0 (leader of pair 0) and CTA 15 (peer of pair 7)
    for cta in (0, 15):
        if cta < t.shape[0]:
            row = t[cta]
            print(f"-- CTA {cta}: " + " ".join(f"{s}:{(row[s] - t0) / 1e3:.1f}" for s in sorted(NAMES) if row[s] > 0))


if __name__ == "__main__":
    main()

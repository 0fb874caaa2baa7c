"""Timeline of attn_tc_kernel from its %globaltimer stamps (debug build path: B200TTS_ATTN_TRACE=<file>).

    B200TTS_GRAPHS=0 B200TTS_ATTN_TRACE=gpurun_out/attn_trace.bin python tools/attn_trace.py [U]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402,F401
from b200tts import capi  # noqa: E402


def main():
    U = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    path = os.environ["B200TTS_ATTN_TRACE"]
    eng = capi.Engine(0)
    rng = np.random.default_rng(0)
    H, N = 16, 1126
    q = (0.35 * rng.standard_normal((2, H, N, 64))).astype(np.float32)
    k = (0.35 * rng.standard_normal((2, H, N, 64))).astype(np.float32)
    v = rng.standard_normal((2, H, N, 64)).astype(np.float32)
    for _ in range(3):
        eng.attention(q, k, v, precision=capi.F16)
    t = np.fromfile(path, dtype=np.uint64).reshape(-1, 64).astype(np.float64)
    t0 = t[t > 0].min()
    rel = lambda x: (x - t0) / 1e3
    print("CTAs", t.shape[0], "kernel span us", rel(t.max()))
    names = {0: "start", 1: "q loaded (mma)", 2: "last pv done (softmax)", 3: "end"}
    for s, n in names.items():
        v_ = t[:, s]; v_ = v_[v_ > 0]
        print(f"{n:28s} min {rel(v_.min()):7.2f} med {rel(np.median(v_)):7.2f} max {rel(v_.max()):7.2f}")
    if os.environ.get("ATTN_TRACE_BRIEF"):
        return
    for j in range(9):
        row = []
        for kk, nm in enumerate(["S issued", "PV issued", "s_full seen", "p_full arrived"]):
            v_ = t[:, 4 + 4 * j + kk]; v_ = v_[v_ > 0]
            row.append(f"{nm} {rel(np.median(v_)):6.2f}")
        print(f"block {j}: " + " | ".join(row))
    for cta in (0, 1, 150, 287):
        r = t[cta]
        print(f"-- CTA {cta}: start {rel(r[0]):.2f} q {rel(r[1]):.2f} " + " ".join(f"[{rel(r[4+4*j]):.1f} {rel(r[4+4*j+2]):.1f} {rel(r[4+4*j+3]):.1f} {rel(r[4+4*j+1]):.1f}]" for j in range(9)) + f" lastpv {rel(r[2]):.2f} end {rel(r[3]):.2f}")


if __name__ == "__main__":
    main()

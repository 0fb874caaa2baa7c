"""Runs the attention kernel several times on the same operands and compares the outputs bit for bit; prints the error against a
float64 reference computed from the 16-bit-rounded operands."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: E402,F401
from b200tts import capi  # noqa: E402


def main():
    import torch
    eng = capi.Engine(0)
    H = 16
    for N in (1126, 300, 4100):
        rng = np.random.default_rng(N)
        q = (0.35 * rng.standard_normal((2, H, N, 64))).astype(np.float32)
        k = (0.35 * rng.standard_normal((2, H, N, 64))).astype(np.float32)
        v = rng.standard_normal((2, H, N, 64)).astype(np.float32)
        q[1], k[1], v[1] = q[0], k[0], v[0]          # twins: the two sequences must come out bit-identical
        for prec, name in ((capi.F16, "f16"), (capi.BF16, "bf16")):
            outs = [eng.attention(q, k, v, precision=prec) for _ in range(6)]
            ndiff = [int((outs[0] != o).sum()) for o in outs[1:]]
            rows = sorted(set(np.argwhere(outs[0] != outs[1])[:, 1].tolist()))[:12] if ndiff[0] else []
            dt = torch.float16 if prec == capi.F16 else torch.bfloat16
            qh, kh, vh = (torch.from_numpy(t).to(dt).double() for t in (q, k, v))
            want = (torch.softmax(qh @ kh.transpose(-1, -2), dim=-1) @ vh).permute(0, 2, 1, 3).reshape(2, N, H * 64).numpy()
            print(f"twins differ in {int((outs[0][0] != outs[0][1]).sum())} elements (max {float(np.abs(outs[0][0] - outs[0][1]).max()):.3e})")
            print(f"N={N} {name} poly={os.environ.get('B200TTS_ATTN_POLY', 'default')}: elements differing from run 0: {ndiff}, first rows {rows}, "
                  f"max err vs f64 {np.abs(outs[0] - want).max():.3e}, max |diff| between runs {max(float(np.abs(outs[0] - o).max()) for o in outs[1:]):.3e}", flush=True)


if __name__ == "__main__":
    main()

"""Debug helper: tcgen05 conv1d vs torch on a few shapes (run on the GPU box)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200tts  # noqa
from b200tts import capi
eng = capi.Engine(0)
cases = [(64, 64, 1, 1, 1, 256, 1), (64, 64, 1, 1, 1, 300, 2), (128, 128, 1, 1, 1, 512, 1), (64, 64, 3, 1, 1, 256, 1), (64, 64, 3, 8, 1, 256, 1),
         (64, 64, 3, 3, 1, 256, 1), (64, 64, 7, 1, 1, 256, 1), (100, 1536, 7, 1, 1, 40, 2), (768, 768, 11, 5, 1, 600, 1), (24, 24, 3, 1, 1, 1000, 2),
         (1024, 1024, 31, 1, 16, 300, 2)]
for (Cin, Cout, k, dil, groups, L, B) in cases:
    rng = np.random.default_rng(0)
    x = rng.standard_normal((B, Cin, L)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin // groups, k)) / np.sqrt(Cin // groups * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    got = eng.conv1d(x, w, b, dilation=dil, groups=groups, precision=capi.BF16)
    xb, wb = torch.from_numpy(x).bfloat16().float(), torch.from_numpy(w).bfloat16().float()
    want = torch.nn.functional.conv1d(xb, wb, torch.from_numpy(b), dilation=dil, padding=(k * dil - dil) // 2, groups=groups).numpy()
    err = np.abs(got - want)
    print(f"Cin={Cin} Cout={Cout} k={k} dil={dil} g={groups} L={L} B={B}: max err {err.max():.4g}  bad rows(t) {np.unique(np.where(err > 1e-2)[2])[:12]}")

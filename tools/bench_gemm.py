"""Micro-benchmark of rowgemm_tc on the hot-path shapes (run on the GPU box). Prints TFLOP/s next to cuBLAS (torch.matmul)
on the equivalent dense problem for orientation."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200tts  # noqa: F401,E402
from b200tts import capi  # noqa: E402

eng = capi.Engine(0)
SHAPES = [
    # name, B, M, N, Cin, taps, dil, groups, epilogue
    ("dit.qkv", 1, 2252, 3072, 1024, 1, 1, 1, 2),
    ("dit.out", 1, 2252, 1024, 1024, 1, 1, 1, 1),
    ("dit.ff1", 1, 2252, 2048, 1024, 1, 1, 1, 2),
    ("dit.ff2", 1, 2252, 1024, 2048, 1, 1, 1, 1),
    ("dit.qkv x4utt", 1, 9008, 3072, 1024, 1, 1, 1, 2),
    ("vgan.s0 k3", 8, 2048, 768, 768, 3, 1, 1, 2),
    ("vgan.s0 k7", 8, 2048, 768, 768, 7, 3, 1, 2),
    ("vgan.s0 k11", 8, 2048, 768, 768, 11, 5, 1, 2),
    ("vgan.s1 k7", 8, 8192, 384, 384, 7, 1, 1, 2),
    ("vgan.s2 k7", 8, 16384, 192, 192, 7, 1, 1, 2),
    ("vgan.s3 k7", 8, 32768, 96, 96, 7, 1, 1, 2),
    ("vgan.s4 k7", 8, 65536, 48, 48, 7, 1, 1, 2),
    ("vgan.s5 k7", 8, 131072, 24, 24, 7, 1, 1, 2),
    ("vgan.s5 k11", 8, 131072, 24, 24, 11, 1, 1, 2),
    # conv2 of a resblock unit: fp32 residual in, fp32 out (epilogue 1)
    ("res.s1 k7", 8, 8192, 384, 384, 7, 1, 1, 1),
    ("res.s2 k7", 8, 16384, 192, 192, 7, 1, 1, 1),
    ("res.s3 k7", 8, 32768, 96, 96, 7, 1, 1, 1),
    ("res.s4 k7", 8, 65536, 48, 48, 7, 1, 1, 1),
    ("res.s5 k7", 8, 131072, 24, 24, 7, 1, 1, 1),
    # the batched DiT (8 utterances per GPU: M = 2*8*1126)
    ("bat.qkv", 1, 18016, 3072, 1024, 1, 1, 1, 2),
    ("bat.out", 1, 18016, 1024, 1024, 1, 1, 1, 1),
    ("bat.ff1", 1, 18016, 2048, 1024, 1, 1, 1, 2),
    ("bat.ff2", 1, 18016, 1024, 2048, 1, 1, 1, 1),
    ("bat.out epi0", 1, 18016, 1024, 1024, 1, 1, 1, 0),
    ("bat.out epi3", 1, 18016, 1024, 1024, 1, 1, 1, 3),
    # epilogue dissection on one thin and one wide stage: 0 fp32 out | 3 fp32 out + residual | 4 bf16 out + residual
    ("epi0.s3 k7", 8, 32768, 96, 96, 7, 1, 1, 0),
    ("epi3.s3 k7", 8, 32768, 96, 96, 7, 1, 1, 3),
    ("epi4.s3 k7", 8, 32768, 96, 96, 7, 1, 1, 4),
    ("epi0.s1 k7", 8, 8192, 384, 384, 7, 1, 1, 0),
    ("epi3.s1 k7", 8, 8192, 384, 384, 7, 1, 1, 3),
    ("epi4.s1 k7", 8, 8192, 384, 384, 7, 1, 1, 4),
]
only = sys.argv[1:] 
for name, B, M, N, Cin, taps, dil, groups, epi in SHAPES:
    if only and not any(o in name for o in only):
        continue
    ms = eng.bench_rowgemm(B, M, N, Cin, taps, dil, groups, epi, iters=20)
    flops = 2.0 * B * M * N * Cin * taps * groups
    # cuBLAS on the same contraction as one dense GEMM (K = taps*Cin)
    a = torch.randn(B * M, taps * Cin, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, taps * Cin, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, w.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms_cb = e0.elapsed_time(e1) / 20
    print(f"{name:14s} B={B} M={M:6d} N={N:4d} Cin={Cin:4d} taps={taps:2d}: ours {ms*1e3:8.1f} us {flops/ms/1e9:7.1f} TF/s | "
          f"cuBLAS dense {ms_cb*1e3:8.1f} us {flops/ms_cb/1e9:7.1f} TF/s", flush=True)

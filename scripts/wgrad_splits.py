"""Sweep split-K for the weight-gradient GEMM shapes of one MPT-125M block (accumulating fp32 output)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from photon_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


T = 32768
for name, M, N in (("qkv", 2304, 768), ("out", 768, 768), ("up", 3072, 768), ("down", 768, 3072)):
    dy = torch.randn(T, M, device="cuda").to(torch.bfloat16)
    x = torch.randn(T, N, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(M, N, device="cuda")
    line = [f"{name:5s} M{M} N{N}"]
    for s in ("auto", 1, 2, 3, 4, 6, 8, 12, 16):
        if s == "auto":
            os.environ.pop("PB_GEMM_SPLITS", None)
        else:
            os.environ["PB_GEMM_SPLITS"] = str(s)
        ms = timeit(lambda: ops.linear_wgrad(dy, x, dw, accumulate=True))
        line.append(f"s={s}:{ms * 1e3:.0f}us/{2 * M * N * T / ms / 1e9:.0f}TF")
    cub = timeit(lambda: torch.addmm(dw, dy.t().float() if False else dy.t().to(torch.bfloat16), x).float() if False else torch.matmul(dy.t(), x))
    line.append(f"cublas(bf16 out):{cub * 1e3:.0f}us")
    print("  ".join(line), flush=True)

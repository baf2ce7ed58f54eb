"""Launch each hot kernel once at flagship shapes so `ncu --set full -k regex:...` can capture it."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from photon_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, S, H, dh = 32, 2048, 12, 64
d = H * dh
T = B * S
which = sys.argv[1] if len(sys.argv) > 1 else "all"
qkv = torch.randn(B, S, 3 * d, device=dev).to(torch.bfloat16)
out = torch.empty(B, S, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device=dev)
if which in ("all", "attn"):
    for _ in range(2):
        ops.attention_fwd(qkv, out, lse, H, 1 / math.sqrt(dh), True)
    dout = torch.randn_like(out)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    for _ in range(2):
        ops.attention_bwd(qkv, out, dout, lse, dqkv, delta, H, 1 / math.sqrt(dh), True)
if which in ("all", "gemm"):
    x = torch.randn(T, d, device=dev).to(torch.bfloat16)
    w = (torch.randn(3 * d, d, device=dev) * 0.02).to(torch.bfloat16)
    y = torch.empty(T, 3 * d, device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(3 * d, device=dev)
    for _ in range(2):
        ops.gemm(x, w, y, bias=bias)                       # qkv forward (K,K)
    dw = torch.zeros(3 * d, d, device=dev)
    for _ in range(2):
        ops.linear_wgrad(y, x, dw)                         # qkv wgrad (MN,MN) split-K reduce-add
if which in ("all", "fp8"):
    # MPT-1B geometry (config #4): per-tensor fp8 forward (qkv) and wgrad, block-scaled MXFP8 forward
    T1, d1 = 16384, 2048
    x = torch.randn(T1, d1, device=dev)
    w = torch.randn(3 * d1, d1, device=dev) * 0.02
    x8 = x.clamp(-400, 400).to(torch.float8_e4m3fn).view(torch.uint8)
    w8 = w.to(torch.float8_e4m3fn).view(torch.uint8)
    y = torch.empty(T1, 3 * d1, device=dev, dtype=torch.bfloat16)
    meta = torch.ones(3, 2, device=dev)
    for _ in range(2):
        ops.gemm_fp8(x8, w8, y, meta, 0, 1, bias=torch.zeros(3 * d1, device=dev))
    dy8 = torch.randn(T1, 3 * d1, device=dev).to(torch.float8_e5m2).view(torch.uint8)
    dw = torch.zeros(3 * d1, d1, device=dev)
    for _ in range(2):
        ops.gemm_fp8(dy8, x8, dw, meta, 0, 1, a_mn=True, b_mn=True, epi=ops.EPI_F32, a_fmt=ops.E5M2, b_fmt=ops.E4M3, accumulate=True)
    xq, xsf = ops.mx_quantize(x.bfloat16())
    wq, wsf = ops.mx_quantize(w.bfloat16(), 2 * ((3 * d1 + 255) // 256))
    for _ in range(2):
        ops.gemm_mxfp8(xq, wq, y, xsf, wsf)
if which in ("all", "round"):
    from photon_b200.parallel.fed_round import NvlFedRound
    from photon_b200.strategy.strategies import FedAdam

    total = 125_440_000 // 4096 * 4096
    fed = NvlFedRound(total, FedAdam(), rank=0, world_size=1, device=dev)
    p0 = torch.randn(total, device=dev)
    fed.set_global(p0)
    for r in range(2):
        fed.begin_round()
        fed.add_client(p0 + 0.01, 2.0)
        fed.finish_round(r + 1)
torch.cuda.synchronize()

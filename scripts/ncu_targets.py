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
torch.cuda.synchronize()

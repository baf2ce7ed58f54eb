"""Forward attention with a share of the exponentials moved from the MUFU pipe to an FMA-pipe polynomial (PB_ATTN_POLY = k: every
k-th pair; 0 = none). One process per setting (the switch is read once): time at both flagship geometries + error against an fp32
reference. Usage: for k in 0 4 3 2; do PB_ATTN_POLY=$k python scripts/attn_poly_sweep.py; done"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, ".")
from photon_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, it=12):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


res = {"PB_ATTN_POLY": int(os.environ.get("PB_ATTN_POLY", "-1"))}
for B, S, H, dh in [(32, 2048, 12, 64), (8, 2048, 16, 128)]:
    d = H * dh
    torch.manual_seed(0)
    qkv = torch.randn(B, S, 3 * d, device=dev).to(torch.bfloat16)
    out = torch.empty(B, S, d, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=dev)
    scale = 1 / math.sqrt(dh)
    res[f"fwd_ms_d{dh}"] = timed(lambda: ops.attention_fwd(qkv, out, lse, H, scale, True))
    q, k, v = qkv[:2].float().view(2, S, 3, H, dh).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(2, S, d)
    res[f"max_abs_err_d{dh}"] = float((out[:2].float() - ref).abs().max())
    res[f"mean_abs_err_d{dh}"] = float((out[:2].float() - ref).abs().mean())
print(json.dumps(res), flush=True)

"""Staged bring-up of the block-scaled MXFP8 GEMM (run on a B200): which part of the scale-factor path is right?"""
import sys

import torch

sys.path.insert(0, ".")
from photon_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def atom_layout(sf_rk: torch.Tensor, rblk: int) -> torch.Tensor:
    """[R, K/32] uint8 -> [K/128][rblk][512] atoms: byte (r%32)*16 + (r//32 % 4)*4 + kg%4 of block (kg//4, r//128)."""
    R, G = sf_rk.shape
    out = torch.zeros(G // 4, rblk, 512, dtype=torch.uint8, device=sf_rk.device)
    r = torch.arange(R, device=dev)[:, None].expand(R, G)
    g = torch.arange(G, device=dev)[None, :].expand(R, G)
    off = (r % 32) * 16 + ((r // 32) % 4) * 4 + (g % 4)
    out[g // 4, r // 128, off] = sf_rk
    return out


def run(M, N, K, sfa_rk, sfb_rk, a8, b8, tag):
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    sfa = atom_layout(sfa_rk, (M + 127) // 128)
    sfb = atom_layout(sfb_rk, 2 * ((N + 255) // 256))
    ops.gemm_mxfp8(a8, b8, out, sfa, sfb)
    torch.cuda.synchronize()
    A = a8.view(torch.float8_e4m3fn).float() * torch.exp2(sfa_rk.float() - 127).repeat_interleave(32, dim=1)
    B = b8.view(torch.float8_e4m3fn).float() * torch.exp2(sfb_rk.float() - 127).repeat_interleave(32, dim=1)
    ref = A @ B.t()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"{tag:55s} rel err {rel:.4g}   ref norm {ref.norm().item():.4g}  out norm {out.float().norm().item():.4g}", flush=True)
    return rel


M, N, K = 256, 512, 256
a8 = (torch.randn(M, K, device=dev)).clamp(-3, 3).to(torch.float8_e4m3fn).view(torch.uint8)
b8 = (torch.randn(N, K, device=dev)).clamp(-3, 3).to(torch.float8_e4m3fn).view(torch.uint8)
one_a = torch.full((M, K // 32), 127, dtype=torch.uint8, device=dev)
one_b = torch.full((N, K // 32), 127, dtype=torch.uint8, device=dev)
run(M, N, K, one_a, one_b, a8, b8, "1. all scales = 1")
row_a = (127 + (torch.arange(M, device=dev) % 5) - 2).to(torch.uint8)[:, None].expand(M, K // 32).contiguous()
run(M, N, K, row_a, one_b, a8, b8, "2. SFA varies per row (same for every K group)")
row_b = (127 + (torch.arange(N, device=dev) % 7) - 3).to(torch.uint8)[:, None].expand(N, K // 32).contiguous()
run(M, N, K, one_a, row_b, a8, b8, "3. SFB varies per row")
kg_a = (127 + (torch.arange(K // 32, device=dev) % 3) - 1).to(torch.uint8)[None, :].expand(M, K // 32).contiguous()
run(M, N, K, kg_a, one_b, a8, b8, "4. SFA varies per K group (same for every row)")
rnd_a = torch.randint(120, 134, (M, K // 32), dtype=torch.uint8, device=dev)
rnd_b = torch.randint(120, 134, (N, K // 32), dtype=torch.uint8, device=dev)
run(M, N, K, rnd_a, rnd_b, a8, b8, "5. random scales on both operands")
# the quantiser + GEMM against the de-quantised reference, odd shape
M, N, K = 640, 776, 1152
x = torch.randn(M, K, device=dev) * torch.exp(torch.randn(M, 1, device=dev))        # rows of very different magnitude
w = torch.randn(N, K, device=dev) * 0.05
xq, xsf = ops.mx_quantize(x.bfloat16())
wq, wsf = ops.mx_quantize(w.bfloat16(), 2 * ((N + 255) // 256))
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
bias = torch.randn(N, device=dev)
ops.gemm_mxfp8(xq, wq, out, xsf, wsf, bias)
torch.cuda.synchronize()


def deq(q, sf, R):
    G = q.shape[1] // 32
    r = torch.arange(R, device=dev)[:, None].expand(R, G)
    g = torch.arange(G, device=dev)[None, :].expand(R, G)
    s = sf[g // 4, r // 128, (r % 32) * 16 + ((r // 32) % 4) * 4 + (g % 4)]
    return q.view(torch.float8_e4m3fn).float() * torch.exp2(s.float() - 127).repeat_interleave(32, dim=1)


ref = deq(xq, xsf, M) @ deq(wq, wsf, N).t() + bias
print("6. mx_quantize + gemm vs de-quantised reference: rel err", ((out.float() - ref).norm() / ref.norm()).item())
full = x.bfloat16().float() @ w.bfloat16().float().t() + bias
print("   vs unquantised product: rel err", ((out.float() - full).norm() / full.norm()).item(), "(MXFP8 quantisation noise)")

"""One MPT-125M local training step (b=32, S=2048) on the engine, for ncu / timing breakdowns.
Use with:  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
           --log-file gpurun_out/launches.csv python scripts/profile_step.py"""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from photon_b200 import ops  # noqa: E402
from photon_b200.models.engine import B200Engine  # noqa: E402
from photon_b200.models.mpt import MPTConfig  # noqa: E402
from photon_b200.train.optim import ADOPT  # noqa: E402

attn = sys.argv[1] if len(sys.argv) > 1 else "b200"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
cfg = MPTConfig()
eng = B200Engine(cfg, dev, "amp_bf16", {"attention": attn})
opt = ADOPT(eng.flat, lr=6e-4, betas=(0.9, 0.9999), eps=1e-6, use_kernel=True, bf16_shadow=eng.bf16_params)
ids = torch.randint(0, 50277, (b, 2048), device=dev)
denom = float(b * 2047)


def step():
    eng.flat.grads.zero_()
    eng.fwd_bwd(ids, denom)
    n = ops.flat_l2_norm(eng.flat.grads)
    opt.step(1.0, torch.clamp(1.0 / (n + 1e-6), max=1.0))


for _ in range(2):
    step()
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


# phase breakdown with CUDA events
ws = eng._workspace(b, 2048)
from photon_b200.models.mpt import shift_labels  # noqa: E402

targets = shift_labels(ids).reshape(-1)
marks = [ev() for _ in range(5)]
eng.flat.grads.zero_()
torch.cuda.synchronize()
marks[0].record()
eng._forward(ids, ws)
marks[1].record()
eng._head(ws, targets, 1.0 / denom, True)
marks[2].record()
eng._backward(ids, ws)
marks[3].record()
n = ops.flat_l2_norm(eng.flat.grads)
opt.step(1.0, torch.clamp(1.0 / (n + 1e-6), max=1.0))
marks[4].record()
torch.cuda.synchronize()
names = ["forward_blocks", "lm_head_ce_fwd_bwd", "backward_blocks", "clip_optimizer"]
phases = {n_: marks[i].elapsed_time(marks[i + 1]) for i, n_ in enumerate(names)}
total = sum(phases.values())
tok = b * 2048
print(json.dumps({"attention": attn, "batch": b, "phases_ms": phases, "step_ms": total, "tokens_per_s": tok / total * 1e3,
                  "launches_per_microbatch": eng.launches_per_microbatch}))
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/phases_{attn}.json").write_text(json.dumps(phases))
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

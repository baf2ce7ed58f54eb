"""Device time of the fused round kernel (and of the DDP all-reduce kernel) over N ranks (torchrun), max over ranks; MPT-125M sized plane.
Knobs through the environment: PB_NVLS=0/1, PB_NVLS_MIN_WORLD, PB_ROUND_WALK=0/1."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from photon_b200.parallel.ddp import NvlGradComm  # noqa: E402
from photon_b200.parallel.fed_round import NvlFedRound  # noqa: E402
from photon_b200.strategy.strategies import FedAdam, FedNesterov  # noqa: E402
from photon_b200.utils.hw import NVLINK_PEER_GBS  # noqa: E402

total = 125_440_000 // 4096 * 4096
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=8):
    ts = []
    for _ in range(n):
        flush.zero_()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


tag = f"N={world} NVLS={os.environ.get('PB_NVLS', '1')} min_world={os.environ.get('PB_NVLS_MIN_WORLD', '4')} walk={os.environ.get('PB_ROUND_WALK', '0')}"
for name, strat in (("nesterov(mu=0)", FedNesterov(1.0, 0.0)), ("fedadam", FedAdam())):
    fed = NvlFedRound(total, strat, rank=rank, world_size=world, device=dev)
    x = torch.randn(total, device=dev)
    fed.set_global(x)
    rnd = [0]

    def one():
        rnd[0] += 1
        fed.finish_round(rnd[0])

    fed.begin_round()
    fed.add_client(x + 0.01, 3.0)
    ms = timed(one)
    roof = total * ((world - 1) / world) * 10 / (NVLINK_PEER_GBS * 1e9) * 1e3 if world > 1 else 0.0
    if rank == 0:
        print(f"[{tag}] fed_round {name}: {ms:.3f} ms (mc={'yes' if fed.arena.mc_ptr('acc') else 'no'}); NVLink roofline {roof:.3f} ms -> {roof / ms:.2f}", flush=True)
    fed.close()
comm = NvlGradComm(total, rank=rank, world_size=world, device=dev)
comm.grads.normal_()
ms = timed(lambda: comm.all_reduce_mean_(comm.grads))
roof = total * 4 * 2 * ((world - 1) / world) / (NVLINK_PEER_GBS * 1e9) * 1e3
g = torch.randn(total, device=dev)
nccl = timed(lambda: dist.all_reduce(g))
if rank == 0:
    print(f"[{tag}] ddp_allreduce: {ms:.3f} ms (roofline {roof:.3f} -> {roof / ms:.2f}); NCCL flat all_reduce of the same bucket: {nccl:.3f} ms", flush=True)
comm.close()
dist.destroy_process_group()

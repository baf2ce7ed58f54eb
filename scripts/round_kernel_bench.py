"""Device time of the fused round kernel alone on one GPU (MPT-125M sized plane, FedAvg and FedAdam), CUDA-event timed, L2 flushed."""
import sys

import torch

sys.path.insert(0, ".")
from photon_b200.parallel.fed_round import NvlFedRound  # noqa: E402
from photon_b200.strategy.strategies import FedAdam, FedAvgEfficient  # noqa: E402
from photon_b200.utils.hw import measured_peaks  # noqa: E402

dev = torch.device("cuda", 0)
total = 125_440_000 // 4096 * 4096
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, strat, bpp in (("fedavg", FedAvgEfficient(1.0), 14), ("fedadam", FedAdam(), 30)):
    fed = NvlFedRound(total, strat, rank=0, world_size=1, device=dev)
    x = torch.randn(total, device=dev)
    fed.set_global(x)
    ts = []
    for it in range(8):
        fed.begin_round()
        fed.add_client(x + 0.01, 3.0)
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fed.finish_round(it + 1)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    roof = total * bpp / measured_peaks()["hbm_bytes_per_s"] * 1e3
    print(f"{name}: {ms:.3f} ms for {total * bpp / 1e9:.2f} GB -> {total * bpp / ms / 1e6:.0f} GB/s, {roof / ms:.2f} of the measured HBM roofline ({roof:.3f} ms)")
    fed.close()

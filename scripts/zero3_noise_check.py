"""How far apart do two runs of the SAME 5-step training end (atomically accumulated gradients make every run slightly different),
and is a fully sharded run (ZeRO-3, one rank) any further from a plain run than a second plain run is? Prints one line per repeat."""
import sys

import torch

sys.path.insert(0, ".")
from photon_b200.data.synthetic import synthetic_batch  # noqa: E402
from photon_b200.models.mpt import MPTConfig  # noqa: E402
from photon_b200.parallel.zero3 import NvlZero3Comm  # noqa: E402
from photon_b200.train.trainer import Trainer  # noqa: E402
from photon_b200.utils.flat import layout_for_model_cfg  # noqa: E402

mc = {"name": "mpt_causal_lm", "d_model": 256, "n_heads": 4, "n_layers": 3, "expansion_ratio": 4, "max_seq_len": 256, "vocab_size": 2048}
cfg = MPTConfig.from_model_cfg(mc)


class Loader:
    def __iter__(self):
        i = 0
        while True:
            ids = torch.from_numpy(synthetic_batch(8, 256, seed=1, start=8 * i) % 2048)
            yield {"input_ids": ids.pin_memory(), "labels": ids}
            i += 1


def make(comm):
    return Trainer(cfg, optimizer_cfg=dict(name="decoupled_adamw", lr=1e-3, betas=[0.9, 0.95], eps=1e-8, weight_decay=1e-4),
                   scheduler_cfg=dict(name="constant_with_warmup", t_warmup="1ba"), train_loader=Loader(), global_train_batch_size=8,
                   device_train_microbatch_size=4, precision="amp_bf16", max_duration="30ba", grad_clip_norm=1.0, device="cuda:0",
                   kernels={"cuda_graph": False}, grad_comm=comm)


def rel(x, y):
    return ((x - y).norm() / y.norm()).item()


for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    a, b = make(None), make(None)
    comm = NvlZero3Comm(layout_for_model_cfg(mc), 3, rank=0, world_size=1, device=torch.device("cuda", 0))
    z = make(comm)
    out = {}
    for steps in (3, 2):
        for t in (a, b, z):
            t.fit(f"{steps}ba")
        n = a.state.timestamp.batch
        out[f"plain_vs_plain@{n}"] = rel(b.state.flat.params, a.state.flat.params)
        out[f"zero3_vs_plain@{n}"] = rel(z.state.flat.full_params(), a.state.flat.params)
    print(rep, {k: f"{v:.2e}" for k, v in out.items()}, flush=True)
    for t in (a, b, z):
        t.close()
    comm.close()

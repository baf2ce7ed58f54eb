"""N1 — DDP gradient all-reduce as one fused NVLink kernel on the flat gradient bucket.

The reference's Composer ``FORCED_SYNC`` issues one NCCL all-reduce per parameter
(148 for MPT-125M) after the last microbatch, un-overlapped (ref:
photon/clients/trainer_utils.py:1714).  Here every rank's flat
gradient buffer is a plane of a :class:`SymmArena`; ``all_reduce_mean_`` launches
``ddp_allreduce_kernel`` (reduce-scatter + all-gather in one pass over peer
pointers, mean folded in, squared-norm partials for clipping as a by-product).
``NcclGradComm`` is the stock-collective baseline with the same interface.
"""
from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist

from photon_b200 import ops
from photon_b200.parallel.symm import SymmArena


class NvlGradComm:
    def __init__(self, total: int, *, rank: int, world_size: int, device: torch.device | int | None = None, group: Any = None) -> None:
        self.total = int(total)
        self.arena = SymmArena({"grads": (total, torch.float32)}, rank=rank, world_size=world_size, device=device, group=group)
        self.norm = torch.zeros(1, dtype=torch.float32, device=torch.device("cuda", self.arena.devices[0]))

    @property
    def grads(self) -> torch.Tensor:
        """The local gradient plane — pass as ``grads_storage`` to :class:`FlatParams`."""
        return self.arena.plane("grads")

    def all_reduce_mean_(self, g: torch.Tensor) -> torch.Tensor:
        ar = self.arena
        if g.data_ptr() != ar.ptrs("grads")[ar.rank]:
            raise ValueError("NvlGradComm reduces its own arena plane; build FlatParams with grads_storage=comm.grads")
        lo, hi = ar.shard(self.total)
        ops.ext().ddp_allreduce(ar.ctl_ptrs(), ar.rank, ar.devices[0], ar.next_epoch(), ar.ptrs("grads"), lo, hi, self.norm)
        return g

    def last_grad_norm(self) -> torch.Tensor:
        """‖mean gradient‖₂ of the last all-reduce (device scalar)."""
        return self.norm[0]

    def close(self) -> None:
        self.arena.close()


class NcclGradComm:
    """Baseline: ONE NCCL all-reduce on the flat bucket (already better than the reference's 148)."""

    def __init__(self, group: Any = None) -> None:
        self.group = group

    def all_reduce_mean_(self, g: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(g, group=self.group)
        return g.div_(dist.get_world_size(self.group))

    def close(self) -> None:
        pass

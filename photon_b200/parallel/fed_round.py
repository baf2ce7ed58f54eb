"""R1 + R2 of a federated round as ONE fused NVLink kernel per GPU.

Host protocol (per round, per GPU):

1. ``begin_round()`` zeroes the local accumulation plane;
2. every client that finishes on this GPU calls ``add_client(params, n_k)`` —
   ``acc += n_k · params`` (one fused axpby; the streaming mean of
   ref: photon/strategy/aggregation.py:57-75 without ever leaving HBM);
3. ``finish_round(t)`` publishes this GPU's Σn_k and launches
   ``fed_round_kernel`` (``csrc/comm.cu``): P2P-pull the peers' slices of
   ``acc``, normalise by the global Σn_k, server optimizer on this GPU's shard of
   (x, m, v), P2P-push the new fp32 slice **and its bf16 cast** into every GPU's
   global planes, L2-norm by-products.  No NCCL, no host copies
   (ref host path: photon/server/fit_utils.py:41-217,
   photon/server/broadcast_utils.py:60-201).

A failed client simply never calls ``add_client`` → its weight is absent from Σn_k,
which is the participation-mask semantics of SURVEY §5.3.
"""
from __future__ import annotations

import math
from typing import Any

import torch

from photon_b200 import ops
from photon_b200.parallel.symm import SymmArena
from photon_b200.strategy.strategies import ServerStrategy

_KIND = {"fedavg": 0, "nesterov": 1, "fedmom": 2, "fedadam": 3, "fedyogi": 4}


class NvlFedRound:
    def __init__(self, total: int, strategy: ServerStrategy, *, rank: int = 0, world_size: int = 1,
                 device: torch.device | int | None = None, group: Any = None, devices: list[int] | None = None,
                 bf16_shadow: bool = True, layout: Any = None) -> None:
        """``layout``: the exchange :class:`FlatLayout` — its tensor boundaries become the kernel's segment table, so the round
        also yields the per-tensor norms (``server/layer/{i}/l2_norm_*``). Without it the whole plane is one segment."""
        if total % 256:
            raise ValueError("flat length must be a multiple of 256 (FlatLayout pads to 4096)")
        self.total, self.strategy = int(total), strategy
        offs = list(layout.offsets) if layout is not None else [0]
        if any(o % 256 for o in offs) or (layout is not None and layout.total != total):
            raise ValueError("layout offsets must be multiples of 256 elements and cover `total`")
        self.n_seg = len(offs)
        self._seg_bounds_host = [o // 256 for o in offs] + [total // 256]
        planes = {"acc": (total, torch.float32), "xg": (total, torch.float32)}
        if bf16_shadow:
            planes["xs"] = (total, torch.bfloat16)
        self.arena = SymmArena(planes, rank=rank, world_size=world_size, device=device, group=group, devices=devices)
        self.n_local = len(self.arena.devices) if self.arena.single else 1
        self.has_shadow = bf16_shadow
        self._wsum = [0.0] * self.n_local
        # server moments live outside the arena (never read by peers) and exist ONLY for the shard a rank owns
        # (8 B/param divided by the number of GPUs: what lets the multi-billion-parameter configs fit); the kernel keeps
        # indexing by global element, so it is handed `shard_ptr - lo`
        self._m: list[torch.Tensor | None] = []
        self._v: list[torch.Tensor | None] = []
        self._seg_bounds: list[torch.Tensor] = []
        self._seg_sums: list[torch.Tensor] = []
        self.last_status = 0
        self._orphans: dict[int, tuple[torch.Tensor | None, torch.Tensor | None]] = {}   # server moments of adopted (dead ranks') shards
        for i in range(self.n_local):
            dev = torch.device("cuda", self.arena.devices[i])
            lo, hi = self.shard_of(i)
            self._seg_bounds.append(torch.tensor(self._seg_bounds_host, dtype=torch.int64, device=dev))
            self._seg_sums.append(torch.zeros(5, self.n_seg, dtype=torch.float64, device=dev))
            self._m.append(torch.zeros(hi - lo, device=dev) if strategy.n_moments >= 1 else None)
            self._v.append(torch.zeros(hi - lo, device=dev) if strategy.n_moments >= 2 else None)

    def shard_of(self, local: int = 0) -> tuple[int, int]:
        """[lo, hi) of the flat index space whose server state lives on local GPU ``local``."""
        rank = local if self.arena.single else self.arena.rank
        return self.arena.shard(self.total, rank)

    # ------------------------------------------------------------------ planes
    def _r(self, local: int) -> int | None:
        return local if self.arena.single else None

    def global_params(self, local: int = 0) -> torch.Tensor:
        return self.arena.plane("xg", self._r(local))

    def global_shadow(self, local: int = 0) -> torch.Tensor | None:
        return self.arena.plane("xs", self._r(local)) if self.has_shadow else None

    def acc(self, local: int = 0) -> torch.Tensor:
        return self.arena.plane("acc", self._r(local))

    def set_global(self, params: torch.Tensor) -> None:
        """Install the initial global model on every local GPU (fp32 + bf16 cast)."""
        for i in range(self.n_local):
            g = self.global_params(i)
            g.copy_(params.to(g.device))
            if self.has_shadow:
                ops.cast_bf16(g, self.global_shadow(i))

    def set_moments(self, m: torch.Tensor | None, v: torch.Tensor | None) -> None:
        """Install full-length server moments (every rank keeps its own slice)."""
        for i in range(self.n_local):
            lo, hi = self.shard_of(i)
            if m is not None and self._m[i] is not None:
                self._m[i].copy_(m[lo:hi].to(self._m[i].device))
            if v is not None and self._v[i] is not None:
                self._v[i].copy_(v[lo:hi].to(self._v[i].device))

    # ------------------------------------------------------------------ round protocol
    def begin_round(self) -> None:
        for i in range(self.n_local):
            self.acc(i).zero_()
            self._wsum[i] = 0.0

    def add_client(self, params: torch.Tensor, weight: float, local: int = 0) -> None:
        if weight <= 0:
            raise ValueError("client weight must be positive")
        ops.axpby_(self.acc(local), params, 1.0, float(weight))
        self._wsum[local] += float(weight)

    def finish_round(self, server_round: int, alive: list[int] | None = None) -> None:
        """Launch the fused reduce + server-opt + broadcast kernel on every local GPU.

        ``alive``: the ranks taking part (default: all). After a rank died the survivors run the SAME kernel over the subset —
        control pages, planes and the reduction simply leave the dead peer out (P2P path; the multicast mapping still includes
        it) — and the slice of the index space the dead rank owned is ADOPTED by a survivor in an extra launch (its server
        moments restart from zero there: that state lived in the dead process)."""
        st, hp, ext, ar = self.strategy, self.strategy.hp, ops.ext(), self.arena
        kind = _KIND[st.kind]
        world = ar.world_size
        ranks = list(range(world)) if alive is None else sorted(int(r) for r in alive)
        degraded = len(ranks) < world
        if degraded and ar.single:
            raise NotImplementedError("masked rounds are a multi-process feature")
        sub = (lambda xs: [xs[r] for r in ranks]) if degraded else (lambda xs: xs)
        acc_mc, xg_mc = (0, 0) if degraded else (ar.mc_ptr("acc"), ar.mc_ptr("xg"))
        shadow = ar.ptrs("xs") if self.has_shadow else []

        def launch(i: int, rank: int, lo: int, hi: int, m: torch.Tensor | None, v: torch.Tensor | None, epoch: int, first: bool = True) -> None:
            # ONE launch per round and GPU: the kernel itself publishes this rank's client-weight sum and zeroes the per-tensor
            # accumulators before its start barrier
            ext.fed_round(sub(ar.ctl_ptrs()), ranks.index(rank), ar.devices[i if ar.single else 0], epoch, sub(ar.ptrs("acc")), sub(ar.ptrs("xg")),
                          sub(shadow) if shadow else [], m.data_ptr() - 4 * lo if m is not None else 0, v.data_ptr() - 4 * lo if v is not None else 0,
                          lo, hi, self.total, kind, st.scaling_factor(), hp.get("lr", 1.0), hp.get("mu", 0.0), hp.get("eta", 0.0),
                          hp.get("beta1", 0.9), hp.get("beta2", 0.99), hp.get("tau", 1e-3), int(server_round), bool(st.sign_compat),
                          acc_mc, xg_mc, self._seg_bounds[i], self._seg_sums[i], float(self._wsum[i]), bool(first))

        epoch = ar.next_epoch()
        for i in range(self.n_local):      # every participant updates the shard it owns
            rank = i if ar.single else ar.rank
            lo, hi = ar.shard(self.total, rank)
            launch(i, rank, lo, hi, self._m[i], self._v[i], epoch)
        if degraded:                        # orphaned shards, one launch each: every survivor takes part, the adopter does the work
            for d in [r for r in range(world) if r not in ranks]:
                adopter = ranks[d % len(ranks)]
                lo, hi = ar.shard(self.total, d)
                epoch = ar.next_epoch()
                if ar.rank == adopter:
                    m, v = self._orphan_moments(d, hi - lo)
                    launch(0, ar.rank, lo, hi, m, v, epoch, first=False)
                else:
                    launch(0, ar.rank, lo, lo, None, None, epoch, first=False)

    def _orphan_moments(self, dead_rank: int, n: int) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        if dead_rank not in self._orphans:
            dev = torch.device("cuda", self.arena.devices[0])
            print(f"[round/nvl] adopting the server shard of dead rank {dead_rank}: its moments restart from zero", flush=True)
            self._orphans[dead_rank] = (torch.zeros(n, device=dev) if self.strategy.n_moments >= 1 else None,
                                        torch.zeros(n, device=dev) if self.strategy.n_moments >= 2 else None)
        return self._orphans[dead_rank]

    def check_status(self) -> int:
        """Sticky status word of the local control page(s) after the round kernel (one 4-byte read, synchronising): 0 = every
        peer took part; bit t = peer t never reached the start barrier within the time-out → the kernel ABORTED without
        touching anything (accumulators, model and moments are as before the launch); bit 8+t = peer t vanished mid-kernel."""
        off = ops.ext().ctl_status_word_offset()
        st = 0
        for i in range(self.n_local):
            st |= int(self.arena.ctl_words(self._r(i))[off].item()) & 0xFFFF
        self.last_status = st
        return st

    def clear_status(self) -> None:
        off = ops.ext().ctl_status_word_offset()
        for i in range(self.n_local):
            self.arena.ctl_words(self._r(i))[off] = 0

    def segment_sq_sums(self, group: Any = None, ctl: Any = None) -> torch.Tensor:
        """float64 [5, n_seg] on the host: per-tensor Σpg², Σa², Σx², Σm², Σv² of the last round, summed over all ranks' shards
        (tiny: off the hot path, called when the metrics are collected)."""
        sums = torch.zeros(5, self.n_seg, dtype=torch.float64)
        for i in range(self.n_local):
            sums += self._seg_sums[i].cpu()
        if ctl is not None:
            return ctl.sum_tensor("segsq", sums)         # host control plane: survives dead peers
        if not self.arena.single and self.arena.world_size > 1:
            import torch.distributed as dist

            t = sums.to(torch.device("cuda", self.arena.devices[0]))
            dist.all_reduce(t, group=group)
            sums = t.cpu()
        return sums

    def round_norms(self, group: Any = None, per_layer: bool = True, ctl: Any = None) -> dict[str, float]:
        """The reference's server norm metrics from the kernel's by-products: ``server/l2_norm_{pseudo_gradient,fedavg_result,
        model,momentum_vector,second_momentum_vector}`` and, per tensor of the layout, ``server/layer/{i}/l2_norm_*``
        (ref: photon/strategy/fedadam.py:333-381)."""
        sums = self.segment_sq_sums(group, ctl)
        names = ["pseudo_gradient", "fedavg_result", "model", "momentum_vector", "second_momentum_vector"][: 3 + self.strategy.n_moments]
        out: dict[str, float] = {}
        for j, n in enumerate(names):
            row = sums[j].tolist()
            out[f"server/l2_norm_{n}"] = math.sqrt(max(sum(row), 0.0))
            if per_layer and self.n_seg > 1:
                for i, q in enumerate(row):
                    out[f"server/layer/{i}/l2_norm_{n}"] = math.sqrt(max(q, 0.0))
        return out

    def moments(self, local: int = 0) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        """This GPU's slices ``[lo, hi)`` of the server moments (see :meth:`shard_of`)."""
        return self._m[local], self._v[local]

    def full_moments(self, local: int = 0) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        """Zero-padded full-length copies of this GPU's slices (sum them over ranks to stitch the state)."""
        lo, hi = self.shard_of(local)
        out = []
        for t in (self._m[local], self._v[local]):
            if t is None:
                out.append(None)
                continue
            full = torch.zeros(self.total, dtype=t.dtype, device=t.device)
            full[lo:hi] = t
            out.append(full)
        return out[0], out[1]

    def add_orphans(self, full: torch.Tensor, which: int) -> torch.Tensor:
        """Add the moments of the shards this rank adopted from dead ranks into a zero-padded full-length plane."""
        for d, mv in self._orphans.items():
            if mv[which] is not None:
                lo, hi = self.arena.shard(self.total, d)
                full[lo:hi] = mv[which]
        return full

    def close(self) -> None:
        self.arena.close()

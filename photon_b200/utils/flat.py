"""Flat parameter storage.

Every trainable tensor of a model lives as a view into ONE contiguous fp32
buffer, laid out in lexicographic name order — the exchange / checkpoint order
of the reference (ref: photon/utils.py:316-317; photon/clients/utils.py:854-857).
The same layout is used for the gradient buffer, the optimizer moments, the
bf16 compute shadow and the symmetric-memory planes the NVLink round kernels
(R1/R2) and the DDP all-reduce (N1) operate on: those kernels take one base
pointer + one length instead of 148 (125M) / 292 (1B) per-tensor launches.

Each tensor starts on an ``align``-element boundary (default 256 → 1 KiB in
fp32, 512 B in bf16) so every weight matrix satisfies TMA's 16-byte global
address rule in both planes.  Padding never leaves this module: the ndarray
codecs below strip it, so payloads/npz files carry exactly the reference's
per-tensor arrays.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np
import torch
import torch.nn as nn


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass(frozen=True)
class FlatLayout:
    names: tuple[str, ...]
    shapes: tuple[tuple[int, ...], ...]
    offsets: tuple[int, ...]
    numels: tuple[int, ...]
    total: int  # padded length in elements
    align: int

    @classmethod
    def build(cls, named_shapes: Iterable[tuple[str, Sequence[int]]], align: int = 256,
              total_multiple: int = 4096) -> "FlatLayout":
        items = sorted(((n, tuple(int(s) for s in shp)) for n, shp in named_shapes), key=lambda kv: kv[0])
        names, shapes, offsets, numels = [], [], [], []
        off = 0
        for n, shp in items:
            ne = int(np.prod(shp)) if len(shp) else 1
            names.append(n), shapes.append(shp), offsets.append(off), numels.append(ne)
            off = _round_up(off + ne, align)
        return cls(tuple(names), tuple(shapes), tuple(offsets), tuple(numels),
                   _round_up(max(off, 1), total_multiple), align)

    @property
    def n_params(self) -> int:
        return int(sum(self.numels))

    def index(self, name: str) -> int:
        return self.names.index(name)

    def view(self, flat: torch.Tensor, i: int | str) -> torch.Tensor:
        i = self.index(i) if isinstance(i, str) else i
        return flat[self.offsets[i]: self.offsets[i] + self.numels[i]].view(self.shapes[i])

    def views(self, flat: torch.Tensor) -> list[torch.Tensor]:
        return [self.view(flat, i) for i in range(len(self.names))]

    def stacked(self, prefixes: Sequence[str]) -> "FlatLayout":
        """This layout repeated once per prefix, plane after plane (``total`` elements each): the exchange layout
        of ``fl.aggregate_momenta`` — ``[params | exp_avg | exp_avg_sq]`` (ref: photon/clients/utils.py:457-468,626-650).
        The first prefix is normally ``""`` so plane 0 keeps the model's names."""
        names, shapes, offsets, numels = [], [], [], []
        for k, pre in enumerate(prefixes):
            names += [pre + n for n in self.names]
            shapes += list(self.shapes)
            offsets += [o + k * self.total for o in self.offsets]
            numels += list(self.numels)
        return FlatLayout(tuple(names), tuple(shapes), tuple(offsets), tuple(numels), self.total * len(prefixes), self.align)

    def segment_table(self) -> torch.Tensor:
        """int64 [n,2] (offset, numel) — consumed by per-layer norm kernels."""
        return torch.tensor(list(zip(self.offsets, self.numels)), dtype=torch.int64)

    # -- ndarray codecs (payload / npz order) -----------------------------------
    def to_ndarrays(self, flat: torch.Tensor) -> list[np.ndarray]:
        host = flat.detach().to("cpu", torch.float32).numpy()
        return [host[o:o + n].reshape(s).copy() for o, n, s in zip(self.offsets, self.numels, self.shapes)]

    def from_ndarrays(self, flat: torch.Tensor, arrays: Sequence[np.ndarray], strict_shapes: bool = True) -> None:
        if len(arrays) != len(self.names):
            raise ValueError(f"expected {len(self.names)} arrays, got {len(arrays)}")
        host = torch.zeros(self.total, dtype=torch.float32)
        hv = host.numpy()
        for a, o, n, s, nm in zip(arrays, self.offsets, self.numels, self.shapes, self.names):
            a = np.asarray(a)
            if strict_shapes and tuple(a.shape) != tuple(s):
                raise ValueError(f"shape mismatch for {nm}: payload {a.shape} vs model {s}")
            if a.size != n:
                raise ValueError(f"size mismatch for {nm}: payload {a.size} vs model {n}")
            hv[o:o + n] = a.reshape(-1).astype(np.float32, copy=False)
        flat.copy_(host.to(flat.device, flat.dtype), non_blocking=False)


class FlatParams:
    """Owns the flat fp32 master buffer (+ grad buffer) of a model and re-points the
    module's ``Parameter.data`` / ``.grad`` at views of them."""

    is_sharded = False      # photon_b200.parallel.zero3.ShardedFlat: ``params`` / ``grads`` are one rank's shard

    def __init__(self, model: nn.Module, align: int = 256, with_grad: bool = True,
                 device: torch.device | str | None = None, params_storage: torch.Tensor | None = None,
                 grads_storage: torch.Tensor | None = None) -> None:
        named = sorted(((n, p) for n, p in model.named_parameters() if p.requires_grad), key=lambda kv: kv[0])
        if not named:
            raise ValueError("model has no trainable parameters")
        dev = torch.device(device) if device is not None else named[0][1].device
        self.layout = FlatLayout.build(((n, p.shape) for n, p in named), align=align)
        # external storage = a plane of the symmetric NVLink arena (so the round / all-reduce
        # kernels can address this buffer on every peer)
        for st in (params_storage, grads_storage):
            if st is not None and (st.numel() < self.layout.total or st.dtype != torch.float32):
                raise ValueError("external flat storage must be float32 with >= layout.total elements")
        self.params = params_storage[: self.layout.total].zero_() if params_storage is not None else \
            torch.zeros(self.layout.total, dtype=torch.float32, device=dev)
        self.grads = (grads_storage[: self.layout.total].zero_() if grads_storage is not None else
                      torch.zeros(self.layout.total, dtype=torch.float32, device=dev)) if with_grad else None
        self._named = named
        with torch.no_grad():
            for i, (_, p) in enumerate(named):
                v = self.layout.view(self.params, i)
                v.copy_(p.detach().to(dev, torch.float32))
                p.data = v
                if with_grad:
                    p.grad = self.layout.view(self.grads, i)

    @property
    def names(self) -> tuple[str, ...]:
        return self.layout.names

    def zero_grad(self) -> None:
        if self.grads is not None:
            self.grads.zero_()
            for i, (_, p) in enumerate(self._named):  # autograd may have replaced .grad
                if p.grad is None or p.grad.data_ptr() != self.layout.view(self.grads, i).data_ptr():
                    p.grad = self.layout.view(self.grads, i)

    def full_params(self) -> torch.Tensor:
        """The whole fp32 vector (the buffer itself here; assembled from the owners under full sharding)."""
        return self.params

    def load_full_params(self, full: torch.Tensor) -> None:
        self.params.copy_(full.to(self.params.device, torch.float32))

    def to_ndarrays(self) -> list[np.ndarray]:
        return self.layout.to_ndarrays(self.params)

    def load_ndarrays(self, arrays: Sequence[np.ndarray]) -> None:
        with torch.no_grad():
            self.layout.from_ndarrays(self.params, arrays)


def layout_for_model_cfg(model_cfg: "object", frozen: list[str] | None = None, unfrozen: list[str] | None = None) -> FlatLayout:
    """Flat layout of an MPT config WITHOUT materialising weights (meta device) — lets callers size
    symmetric-memory planes before the Trainer exists."""
    from photon_b200.models.mpt import MPTConfig, MPTForCausalLM

    mc = model_cfg if isinstance(model_cfg, MPTConfig) else MPTConfig.from_model_cfg(dict(model_cfg))
    model = MPTForCausalLM(mc, device="meta", init=False)
    names = [(n, tuple(p.shape)) for n, p in model.named_parameters()
             if not (frozen and n in frozen) and not (unfrozen and n not in unfrozen)]
    return FlatLayout.build(names)

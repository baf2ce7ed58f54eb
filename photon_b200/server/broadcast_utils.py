"""Server→nodes broadcast (R2). In the reference this is a QUERY message per node plus a
side-channel payload and a busy-poll for ``{"broadcast": {"status": "OK"}}`` acks (ref:
photon/server/broadcast_utils.py:60-201). Here the new global model already sits in every
rank's global planes when the round transport finishes (the ``nvl`` kernel pushes it over
NVLink; the other transports end with an H2D), so "broadcast" reduces to installing the
initial / restored model and collecting the acks."""
from __future__ import annotations

import time
from typing import Any

import torch
import torch.distributed as dist


def broadcast_parameters_to_nodes(runtime: Any, params: torch.Tensor | None, momentum: torch.Tensor | None = None,
                                  second: torch.Tensor | None = None, src: int = 0) -> dict[str, Any]:
    """Install ``params`` (rank ``src``'s copy wins) as the global model on every node."""
    t0 = time.time()
    dev = runtime.device
    total = runtime.layout.total
    bufs = []
    for t, need in ((params, True), (momentum, runtime.strategy.n_moments >= 1), (second, runtime.strategy.n_moments >= 2)):
        if not need:
            bufs.append(None)
            continue
        b = t.to(dev, torch.float32).clone() if t is not None else torch.zeros(total, device=dev)
        if runtime.world_size > 1 and dist.is_initialized():
            dist.broadcast(b, src=src, group=runtime.group)
        bufs.append(b)
    runtime.round_backend.set_global(bufs[0], bufs[1], bufs[2])
    acks = [{"broadcast": {"status": "OK"}}] * runtime.world_size
    return {"acks": acks, "server/broadcast_time": time.time() - t0}


def parameters_to_broadcast_recordset(parameters: Any, *, keep_input: bool = True, node_id: int = 0) -> Any:
    """The QUERY message that carries a broadcast (ref: broadcast_utils.py:28-57 builds a RecordSet with the parameters and a
    ``{"action": "set_parameters"}`` config): ``parameters`` is a flat tensor, an ndarray list or already a ``ParamHandle``
    (side-channel locator). ``keep_input=False`` drops the caller's reference to an inline payload once it is in the message, like
    the reference's conversion does."""
    from photon_b200.messages import Message, ParamHandle

    handle = parameters if isinstance(parameters, ParamHandle) else ParamHandle("inline", parameters)
    msg = Message("query", {"type": "broadcast_parameters", "action": "set_parameters", "parameters": handle}, node_id=node_id)
    if not keep_input and isinstance(parameters, list):
        parameters.clear()
    return msg

"""Round-0 initialisation and resume (ref: photon/server/init_utils.py:128-287)."""
from __future__ import annotations

from typing import Any

import torch

from photon_b200.checkpoint.store import CheckpointStore
from photon_b200.messages import ClientState, decode_client_states, encode_client_states
from photon_b200.server.broadcast_utils import broadcast_parameters_to_nodes
from photon_b200.strategy.strategies import MOMENTUM_KEY, SECOND_MOMENTUM_KEY, SERVER_PARAMETERS_KEY
from photon_b200.wandb_history import WandbHistory


def server_state_dict(runtime: Any, history: WandbHistory, time_offset: float) -> dict[str, Any]:
    """Contents of ``state.bin`` (ref: photon/server/s3_utils.py:374-389)."""
    return {"history": history, "time_offset": float(time_offset), "client_state": encode_client_states(runtime.client_states),
            "server_steps_cumulative": int(runtime.server_steps_cumulative)}


def get_centralized_run_parameters(cfg: Any, layout: Any) -> torch.Tensor | None:
    """Warm-start the federation from a centralised run's ``{uuid}-{batches}-checkpoint.npz``
    (``photon.restore_cent_run_uuid`` / ``restore_cent_run_batches``; ref: init_utils.py:43-125 — which is
    broken there: it dereferences non-existent ``llm_config.fl`` keys)."""
    import glob
    import os

    ph = cfg["photon"]
    uuid = ph.get("restore_cent_run_uuid")
    if not uuid:
        return None
    root = ph.get("saving_path") or os.environ.get("PHOTON_SAVE_PATH", ".")
    batches = ph.get("restore_cent_run_batches")
    pattern = f"{uuid}-{batches if batches is not None else '*'}-checkpoint.npz"
    cands = sorted(glob.glob(os.path.join(root, "**", pattern), recursive=True) + glob.glob(pattern),
                   key=lambda p: int(p.rsplit("-", 2)[-2]))
    if not cands:
        raise FileNotFoundError(f"no centralised checkpoint matching {pattern} under {root}")
    from photon_b200.utils.core import load_model_parameters_from_file

    flat = torch.zeros(layout.total, dtype=torch.float32)
    layout.from_ndarrays(flat, load_model_parameters_from_file(cands[-1])[: len(layout.names)])
    return flat


def initialize_round(runtime: Any, store: CheckpointStore | None, history: WandbHistory) -> int:
    """Zero state, fresh ``ClientState`` per client, initial parameters, zero momenta, round-0
    checkpoint (ref: init_utils.py:128-223). Returns the round to start from (0)."""
    n = int(runtime.cfg["fl"]["n_total_clients"])
    runtime.client_states = {c: ClientState() for c in range(n)}
    runtime.server_steps_cumulative = 0
    params = runtime.initial_parameters()
    if runtime.rank == 0:
        model_layout = getattr(runtime, "model_layout", None) or runtime.layout
        cent = get_centralized_run_parameters(runtime.cfg, model_layout)
        if cent is not None:
            params[: model_layout.total] = cent  # any momenta planes stay zero
    broadcast_parameters_to_nodes(runtime, params)
    if store is not None and runtime.rank == 0:
        store.upload_server_checkpoint(str(runtime.cfg["run_uuid"]), 0, layout=runtime.layout, tensors=runtime.state_tensors(),
                                       state=server_state_dict(runtime, history, 0.0))
    elif store is not None:
        runtime.state_tensors()  # collective inside (moment gather) must be entered by every rank
    return 0


def resume_from_round(runtime: Any, store: CheckpointStore, run_uuid: str, server_round: int) -> tuple[WandbHistory, float]:
    """Load a complete round, replay the client-sampling RNG and re-install model + momenta on
    every node (ref: init_utils.py:226-287, photon/server_app.py:188-192)."""
    keys = runtime.strategy.state_keys
    history, time_offset = WandbHistory(bool(runtime.cfg.get("use_wandb", False))), 0.0
    tensors: dict[str, torch.Tensor] = {}
    if runtime.rank == 0:
        tensors, state = store.download_server_checkpoint(run_uuid, server_round, layout=runtime.layout, state_keys=keys)
        runtime.client_states = decode_client_states(state["client_state"])
        runtime.server_steps_cumulative = int(state["server_steps_cumulative"])
        history, time_offset = state.get("history", history), float(state.get("time_offset", 0.0))
    if runtime.world_size > 1:
        import torch.distributed as dist

        box = [encode_client_states(runtime.client_states), runtime.server_steps_cumulative]
        dist.broadcast_object_list(box, src=0, group=runtime.group)
        runtime.client_states, runtime.server_steps_cumulative = decode_client_states(box[0]), int(box[1])
    runtime.replay_sampling(server_round)
    broadcast_parameters_to_nodes(runtime, tensors.get(SERVER_PARAMETERS_KEY), tensors.get(MOMENTUM_KEY), tensors.get(SECOND_MOMENTUM_KEY))
    return history, time_offset

"""Federated server entry point: the round loop.

Same control flow as the reference's Flower ``ServerApp`` (ref: photon/server_app.py:85-424):
init | resume | restore → broadcast → optional eval → for each round: health check, seeded
client sampling, ``fit_round`` (which ends with the fused aggregate+broadcast), periodic
evaluation, server checkpoint, timing metrics, optional per-round cleanup.  It runs SPMD:
launch it on every GPU with ``torchrun`` (or as one process for 1 GPU / CPU); rank 0 owns the
history and the checkpoint store.

    PHOTON_SAVE_PATH=... python -m photon_b200.hydra_resolver <overrides>
    PHOTON_SAVE_PATH=... torchrun --nproc-per-node 8 -m photon_b200.server_app
"""
from __future__ import annotations

import os
import time
from pathlib import Path
from typing import Any

import torch.distributed as dist

from photon_b200.checkpoint.store import CheckpointStore
from photon_b200.clients.trainer_utils import initialize_dist, pick_device
from photon_b200.config import load_config
from photon_b200.federation import FederationRuntime
from photon_b200.server.evaluate_utils import evaluate_round
from photon_b200.server.fit_utils import fit_round
from photon_b200.server.init_utils import initialize_round, resume_from_round, server_state_dict
from photon_b200.server.server_util import wait_for_nodes_to_connect
from photon_b200.utils.core import wandb_init
from photon_b200.utils.trace import tracer
from photon_b200.wandb_history import WandbHistory


def _store_for(cfg: Any) -> CheckpointStore | None:
    if not (cfg["photon"]["checkpoint"] or cfg["photon"]["comm_stack"].get("s3")):
        return None
    root = cfg["photon"].get("saving_path") or os.environ.get("PHOTON_SAVE_PATH", ".")
    from photon_b200.utils.objstore import remote_store_from_cfg

    return CheckpointStore(root, str(cfg["s3_comm_config"]["bucket_name"]), remote=remote_store_from_cfg(cfg))


def run_server(cfg: Any, runtime: FederationRuntime | None = None, n_rounds: int | None = None) -> WandbHistory:
    own = runtime is None
    if runtime is None and str(cfg["photon"].get("topology", "spmd")) == "nodes":
        from photon_b200.server.fleet import NodeFleetRuntime

        runtime = NodeFleetRuntime(cfg)
    if runtime is None:
        device = pick_device(int(os.environ.get("LOCAL_RANK", "0")))
        rank, world = initialize_dist(device)
        runtime = FederationRuntime(cfg, device=device, rank=rank, world_size=world)
    runtime.build()
    fl, ph, run_uuid = cfg["fl"], cfg["photon"], str(cfg["run_uuid"])
    store = _store_for(cfg)
    history = WandbHistory(bool(cfg.get("use_wandb", False)))
    wandb_run = wandb_init(bool(cfg.get("use_wandb", False)) and runtime.rank == 0, **dict((cfg.get("wandb") or {}).get("setup") or {}))
    t_zero = time.time()

    # ---- init / resume / restore from another run
    start_round, time_offset = 0, 0.0
    resume = None
    if store is not None:
        if ph.get("restore_run_uuid") and runtime.rank == 0:
            store.import_checkpoints(cfg, runtime.strategy.state_keys)
        if runtime.world_size > 1:
            dist.barrier(group=runtime.group)
        resume = store.interpret_resume_round(run_uuid, ph.get("resume_round"), runtime.strategy.state_keys)
    t_bcast = time.time()
    if resume is not None and resume > 0:
        history, time_offset = resume_from_round(runtime, store, run_uuid, resume)
        start_round = resume
    else:
        initialize_round(runtime, store, history)
    if runtime.rank == 0:   # same keys as the reference's dashboards (ref: server_app.py:271-274)
        history.add_metrics_centralized(start_round + 1, {"server/broadcast_pre_time": time.time() - t_bcast})

    wait_for_nodes_to_connect(runtime.n_nodes, runtime.node_ids, poll_s=0.0)
    eval_period = fl.get("eval_period")
    if eval_period and start_round == 0:
        loss, em = evaluate_round(runtime, 0)
        if runtime.rank == 0 and loss is not None:
            history.add_loss_distributed(0, loss)
            history.add_metrics_distributed(0, em)

    total = int(fl["n_rounds"]) if n_rounds is None else start_round + n_rounds
    for server_round in range(start_round + 1, total + 1):
        t_round = time.time()
        t_chk = time.time()
        node_ids = runtime.node_ids()                              # health check (ref: server_app.py:285)
        first_check = time.time() - t_chk
        sampled = runtime.sample_clients()
        with tracer().span("fit_round", cat="server", server_round=server_round, clients=str(sampled)):
            metrics = fit_round(runtime, server_round, sampled)
        metrics["server/n_nodes"] = len(node_ids)
        if runtime.rank == 0:
            history.add_metrics_distributed_fit(server_round, metrics)
        if eval_period and server_round % int(eval_period) == 0:
            with tracer().span("evaluate_round", cat="server", server_round=server_round):
                loss, em = evaluate_round(runtime, server_round)
            if runtime.rank == 0 and loss is not None:
                history.add_loss_distributed(server_round, loss)
                history.add_metrics_distributed(server_round, em)
        if store is not None:
            tensors = runtime.state_tensors()                      # collective when moments are sharded
            if runtime.rank == 0:
                with tracer().span("server_checkpoint", cat="server", server_round=server_round):
                    bg = bool(ph.get("async_checkpoint", True))    # snapshot now, write on the store's writer thread
                    store.upload_server_checkpoint(run_uuid, server_round, layout=runtime.layout, tensors=tensors, background=bg,
                                                   state=server_state_dict(runtime, history, time_offset + time.time() - t_zero))
                    if cfg.get("cleanup_checkpoints_per_round"):
                        if bg:
                            store.submit(lambda: store.cleanup_checkpoints(run_uuid, per_round=True))   # after the write it follows
                        else:
                            store.cleanup_checkpoints(run_uuid, per_round=True)
        t_chk = time.time()
        n_after = len(runtime.node_ids())                          # second liveness check after the round's broadcast (ref: :346)
        if runtime.rank == 0:
            history.add_metrics_centralized(server_round, {
                "server/round_time": time.time() - t_round, "server/first_check_nm_time": first_check,
                "server/second_check_nm_time": time.time() - t_chk, "server/n_nodes_after_round": n_after,
                # reduce + server optimizer + broadcast of the round (the reference times only its broadcast here)
                "server/broadcast_post_time": float(runtime.timings.get("aggregate_broadcast_host_s", runtime.timings.get("server/broadcast_time", 0.0))),
                # parameter bytes that crossed the network this round (node-fleet topology with remote nodes)
                **{k: float(v) for k, v in runtime.timings.items() if k.startswith("comm/")}})
    if store is not None and runtime.rank == 0:
        store.wait()                                               # every round's checkpoint is on disk before the run reports done
    if store is not None and cfg.get("cleanup_checkpoints") and runtime.rank == 0:
        store.cleanup_checkpoints(run_uuid)
    if wandb_run is not None:
        wandb_run.finish()
    tracer().flush()
    if own:
        runtime.close()
    return history


def dump_history(history: WandbHistory, path: str | os.PathLike) -> Path:
    """The run's round-level metrics as plain JSON next to ``config.yaml`` (the reference only has them in wandb or inside the
    pickled ``state.bin``): ``{store: {key: [[round, value], …]}}`` for the five History stores."""
    import json

    def plain(v: Any) -> Any:
        try:
            return float(v) if not isinstance(v, (str, bool, int)) else v
        except (TypeError, ValueError):
            return str(v)

    out = {"losses_distributed": [[r, plain(v)] for r, v in history.losses_distributed],
           "losses_centralized": [[r, plain(v)] for r, v in history.losses_centralized]}
    for name in ("metrics_distributed_fit", "metrics_distributed", "metrics_centralized"):
        out[name] = {k: [[r, plain(v)] for r, v in vals] for k, vals in getattr(history, name).items()}
    p = Path(path)
    p.write_text(json.dumps(out, indent=1))
    return p


def main() -> None:
    save_path = os.environ.get("PHOTON_SAVE_PATH")
    if not save_path:
        raise SystemExit("PHOTON_SAVE_PATH must point at the directory holding config.yaml")
    cfg = load_config(Path(save_path) / "config.yaml")
    hist = run_server(cfg)
    if int(os.environ.get("RANK", "0")) == 0:
        dump_history(hist, Path(save_path) / "history.json")
        last = {k: v[-1][1] for k, v in hist.metrics_distributed_fit.items() if "/layer/" not in k}
        print("[server] done. last round fit metrics:", {k: (round(v, 5) if isinstance(v, float) else v) for k, v in last.items()})


if __name__ == "__main__":
    main()

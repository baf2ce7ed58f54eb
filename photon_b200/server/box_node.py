"""A whole multi-GPU box as ONE node of a cross-host federation: ``python -m photon_b200.node --server host:port --spmd``.

Inside the box the clients are trained by the SPMD runtime (:class:`photon_b200.federation.FederationRuntime`): one process per GPU,
the box's share of the round is a work queue over its GPUs, and the clients' models meet in the fused NVLink round kernel — the
weighted sum, its mean and the bf16 cast never leave the GPUs. Towards the server the box looks like any other node of the gRPC
fleet (:mod:`photon_b200.server.grpc_fleet`) that pre-aggregates: it accepts up to ``capacity`` (= GPUs) clients per task, answers
with metrics only (``ParamHandle("deferred")``) and hands over ONE weighted mean per round when the server collects. Two levels of
the same reduction: NVLink / NVSwitch inside a box, the network between boxes — the shape of the reference's cross-silo setting
(1–10 Gbps links) with B200 boxes as silos.

Process layout: every rank of the box runs :func:`serve_box` (launched by ``torchrun`` or ``photon_b200.launch``). Rank 0 owns the
gRPC connection; what it receives is replayed on the other ranks through the runtime's host control plane (``ControlPlane.broadcast``,
a TCPStore — no collective is blocked while the box waits for work), the round's parameters reach the peers through a file in
``/dev/shm`` (no collective either: a rank that died earlier cannot hold up the install).
"""
from __future__ import annotations

import copy
import os
from typing import Any

import numpy as np
import torch

from photon_b200.messages import Code, EvaluateRes, FitRes, Message, ParamHandle, Status, decode_client_states


def box_config(cfg: Any) -> Any:
    """The box's own view of the run: same model / data / client settings, but its "server" only averages (FedAvg, η = 1): what the
    box hands to the real server is the weighted MEAN of the clients it trained, the server optimizer runs once, on the server."""
    c = copy.deepcopy(cfg)
    c["fl"]["strategy_name"] = "fedavg"
    c["fl"]["strategy_kwargs"] = {}
    c["fl"]["eval_period"] = None
    c["fl"]["fault_injection"] = None
    c["photon"]["checkpoint"] = False
    c["photon"]["topology"] = "spmd"
    if torch.cuda.is_available():
        c["photon"]["comm_stack"] = {"s3": False, "shm": False, "ray": False, "nvl": True}
    else:
        c["photon"]["comm_stack"] = {"s3": False, "shm": False, "ray": True, "nvl": False}
    return c


class BoxApp:
    """``ClientApp`` interface (``handle`` / ``alive`` / ``start`` / ``shutdown``) on top of the SPMD runtime of this box."""

    def __init__(self, cfg: Any, node_id: int, runtime: Any) -> None:
        self.cfg, self.node_id, self.rt = cfg, node_id, runtime
        self._agg: list[np.ndarray] | None = None
        self._agg_w = 0
        self._agg_round = -1
        self._op = 0
        self._global: torch.Tensor | None = None      # the server's model of this round (every batch of the round starts from it)
        self._shm_file: str | None = None
        if runtime.ctl is not None:
            # a box waits for work for as long as the server has none: "rank 0 has not spoken for two hours" is not an error here
            # (a DEAD rank 0 is still noticed — its heartbeat goes stale — and rank 0 keeps ticking while it polls the server)
            runtime.ctl.exchange_timeout_s = 10 * 365 * 86400.0

    # ------------------------------------------------------------------ lifecycle
    def start(self) -> None:
        pass

    def describe(self) -> str:
        return f"SPMD box, {self.rt.world_size} rank(s), transport {self.rt.round_backend.name}"

    def alive(self) -> bool:
        return True

    def tick(self) -> None:
        """Called by the serving loop between polls: waiting for the server IS this rank's progress (the control plane stops the
        heartbeat of a rank whose main thread made none for ``photon.progress_timeout_s``)."""
        if self.rt.ctl is not None:
            self.rt.ctl.tick()

    def shutdown(self) -> None:
        self._replay({"op": "stop"})
        if self._shm_file and os.path.exists(self._shm_file):
            os.unlink(self._shm_file)
        self.rt.close()

    # ------------------------------------------------------------------ rank 0 -> followers
    def _replay(self, op: dict[str, Any]) -> None:
        if self.rt.ctl is not None:
            self._op += 1
            self.rt.ctl.broadcast(f"box_op/{self._op}", op)

    def _install(self, arrays: list[np.ndarray] | None, path: str | None = None) -> None:
        """The round's global model on every rank of the box. Rank 0 holds ``arrays``; the other ranks read the flat vector rank 0 left
        in ``/dev/shm`` (same machine) — no collective, so a rank of the box that died earlier cannot block the ones that are left."""
        lay = self.rt.layout
        if arrays is not None:
            flat = torch.zeros(lay.total, dtype=torch.float32)
            lay.from_ndarrays(flat, arrays)
            if self.rt.ctl is not None:
                new = f"/dev/shm/photon_box_{os.getpid()}_{self._op + 1}.pt"
                torch.save(flat, new)
                self._replay({"op": "install", "path": new})
                if self._shm_file and os.path.exists(self._shm_file):
                    os.unlink(self._shm_file)          # everybody has long finished reading the previous round's vector
                self._shm_file = new
        else:
            flat = torch.load(path, map_location="cpu")
        self.rt.round_backend.set_global(flat.to(self.rt.device), None, None)
        self._global = self.rt.round_backend.global_params().detach().clone()

    def _train(self, server_round: int, cids: list[int], client_state: Any, steps: int) -> tuple[list[FitRes], int]:
        rt = self.rt
        if self._global is None:
            raise RuntimeError("the box was asked to train before it received the round's parameters")
        # a box may get several batches in one round: each starts from the SERVER's model, not from the previous batch's mean
        rt.round_backend.set_global(self._global.clone(), None, None)
        rt.client_states = {int(k): v for k, v in decode_client_states(client_state).items()} if client_state else rt.client_states
        rt.server_steps_cumulative = int(steps)
        results = rt.run_clients_fit(server_round, cids)
        everything = rt.gather_results(results, cids)
        ok = [r for r in everything if r.status.code == Code.OK]
        weight = int(sum(r.num_examples for r in ok))
        if ok:
            rt.finish_round(server_round)        # FedAvg with eta = 1: the global planes now hold the weighted mean of THESE clients
        else:
            rt.abort_round()
        return everything, weight

    # ------------------------------------------------------------------ handlers (rank 0)
    def handle(self, msg: Message) -> Message:
        try:
            if msg.kind == "train":
                return self._handle_train(msg)
            if msg.kind == "evaluate":
                return self._handle_evaluate(msg)
            kind = (msg.content or {}).get("type")
            if kind == "broadcast_parameters":
                payload = msg.content["parameters"]
                if isinstance(payload, ParamHandle):
                    from photon_b200.server.s3_utils import replace_parameters_in_recordset_with_remote

                    payload = replace_parameters_in_recordset_with_remote(payload).data
                self._install(list(payload))
                return Message("query", {"broadcast": {"status": "OK"}}, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)
            if kind == "collect_aggregate":
                return self._collect(msg)
            if kind == "free_resources":
                return Message("query", {"free_resources": {"status": "OK"}}, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)
            return Message("query", None, node_id=self.node_id, error=f"unknown query type {kind!r}", reply_to=msg.msg_id)
        except Exception as e:  # noqa: BLE001 - a node never crashes the federation; it replies with an error
            bad: Any = [FitRes(Status(Code.FAILED, repr(e)), None, 0, {}, c) for c in (getattr(msg, "per_client", {}) or [None])] \
                if msg.kind == "train" else (EvaluateRes(Status(Code.FAILED, repr(e)), 0.0, 0, {}) if msg.kind == "evaluate" else None)
            return Message(msg.kind, bad, node_id=self.node_id, error=repr(e), reply_to=msg.msg_id)

    def _handle_train(self, msg: Message) -> Message:
        conf = msg.content.config
        server_round, steps = int(conf["server_round"]), int(conf.get("server_steps_cumulative", 0) or 0)
        cids = [int(c) for c in getattr(msg, "per_client", {})]
        self._replay({"op": "train", "round": server_round, "cids": cids, "client_state": conf.get("client_state"), "steps": steps})
        results, weight = self._train(server_round, cids, conf.get("client_state"), steps)
        if weight > 0:
            mean = self.rt.round_backend.global_params().detach().to("cpu", torch.float32)
            arrays = [np.asarray(a, dtype=np.float64) * weight for a in self.rt.layout.to_ndarrays(mean)]
            if self._agg_round != server_round:
                self._agg, self._agg_w, self._agg_round = None, 0, server_round
            self._agg = arrays if self._agg is None else [x + y for x, y in zip(self._agg, arrays)]
            self._agg_w += weight
        out = [FitRes(r.status, ParamHandle("deferred", None, {"node_id": self.node_id}) if r.status.code == Code.OK else None,
                      r.num_examples, r.metrics, r.cid) for r in results]
        return Message("train", out, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)

    def _collect(self, msg: Message) -> Message:
        want = int((msg.content or {}).get("server_round", self._agg_round))
        if self._agg is None or self._agg_round != want or self._agg_w <= 0:
            body: Any = {"aggregate": None, "num_examples": 0}
        else:
            body = {"aggregate": ParamHandle("inline", [(a / self._agg_w).astype(np.float32) for a in self._agg]), "num_examples": self._agg_w}
        self._agg, self._agg_w = None, 0
        return Message("query", body, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)

    def _handle_evaluate(self, msg: Message) -> Message:
        conf = msg.content.config
        server_round = int(conf["server_round"])
        cids = [int(c) for c in getattr(msg, "per_client", {})] or [0]
        self._replay({"op": "evaluate", "round": server_round, "cids": cids})
        res = self.rt.run_clients_evaluate(server_round, cids)
        first = res[0] if res else EvaluateRes(Status(Code.FAILED, "no result"), 0.0, 0, {})
        return Message("evaluate", first, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)

    # ------------------------------------------------------------------ followers (ranks 1..)
    def follow(self) -> None:
        """Replay rank 0's operations until it says stop."""
        while True:
            self._op += 1
            op = self.rt.ctl.broadcast(f"box_op/{self._op}", None)
            if op["op"] == "stop":
                self.rt.close()
                return
            if op["op"] == "install":
                self._install(None, op["path"])
            elif op["op"] == "train":
                self._train(op["round"], op["cids"], op["client_state"], op["steps"])
            elif op["op"] == "evaluate":
                self.rt.run_clients_evaluate(op["round"], op["cids"])


def serve_box(server_address: str, *, token: str | None = None, tls_ca: str | None = None, max_idle_s: float | None = None) -> int:
    """Entry point of EVERY rank of the box (``RANK`` / ``WORLD_SIZE`` / ``MASTER_*`` from torchrun or ``photon_b200.launch``)."""
    import torch.distributed as dist

    from photon_b200.clients.trainer_utils import initialize_dist, pick_device
    from photon_b200.federation import FederationRuntime
    from photon_b200.server.grpc_fleet import serve_node

    device = pick_device(int(os.environ.get("LOCAL_RANK", "0")))
    rank, world = initialize_dist(device)

    def build(cfg: Any, node_id: int) -> BoxApp:
        holder = [cfg, node_id]
        if world > 1:
            dist.broadcast_object_list(holder, src=0)        # the followers learn the run's config and the node id from rank 0
        rt = FederationRuntime(box_config(holder[0]), device=device, rank=rank, world_size=world)
        rt.build()
        if world > 1 and rt.ctl is None:
            raise RuntimeError("a multi-rank box needs the host control plane (photon.control_plane=store) to replay rank 0's tasks")
        return BoxApp(holder[0], int(holder[1]), rt)

    if rank == 0:
        return serve_node(server_address, token=token, tls_ca=tls_ca, max_idle_s=max_idle_s, app_factory=build,
                          info={"capacity": world, "kind": "spmd-box"})
    app = build(None, -1)
    app.follow()
    return app.node_id

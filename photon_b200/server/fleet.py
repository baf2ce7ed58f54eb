"""Node-fleet runtime: the reference's process topology, in one box.

``photon.topology: nodes`` runs the federation the way the reference deploys it:

    server loop ──messages──▶ N × ClientApp ──▶ NodeManagerApp ──queues/shm──▶ one Worker process per device

* the server owns no trainer: it holds the global model + strategy on the host, samples clients, and hands them to
  nodes through the **work queue** (first ``n_nodes`` clients go out at once, every reply frees its node for the next
  sampled client — virtual-client multiplexing; ref: photon/server/server_util.py:65-202);
* every round starts with the **broadcast** QUERY to every node — one payload parked once on the parameter side
  channel, one control message per node, acknowledged with ``{"broadcast": {"status": "OK"}}``
  (ref: photon/server/broadcast_utils.py:60-201);
* replies are folded into the running weighted sum **as they arrive** (streaming aggregation, one client payload alive
  at a time; ref: photon/server/fit_utils.py:41-217, photon/strategy/aggregation.py:57-118);
* all devices of a node collaborate on ONE client at a time (DDP / ZeRO over the node's workers), failed workers are
  respawned and the client retried (ref: photon/node_manager/node_manager_app.py:405-592).

The SPMD runtime (:mod:`photon_b200.federation`) is the fast path — one process per GPU and the fused NVLink round kernel.
Nodes on OTHER machines join through the gRPC fleet link (``photon.fleet.address`` + ``photon.fleet.n_remote_nodes``,
:mod:`photon_b200.server.grpc_fleet`; start ``python -m photon_b200.node --server host:port`` on each machine): they are scheduled
through the same work queue as the in-process ones, get their parameters through the S3 bucket (or inline in the gRPC message
when no object store is configured) and drop out of ``node_ids()`` when they stop polling.

This one exists for parity with the reference's deployment shape (persistent worker processes that can be recycled every
``photon.refresh_period`` rounds, host-side server) and as the oracle the SPMD path is compared against.
"""
from __future__ import annotations

import time
import uuid
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any

import torch

from photon_b200.client_app import ClientApp
from photon_b200.clients.utils import get_raw_model_parameters
from photon_b200.federation import FederationRuntime
from photon_b200.messages import Code, EvaluateRes, FitRes, Message, ParamHandle, Status
from photon_b200.server.round_backends import CollectiveRoundBackend
from photon_b200.server.s3_utils import release_remote_parameters, replace_remote_with_parameters_in_recordset
from photon_b200.server.server_util import ClientScheduler, fit_or_evaluate_ins, wait_for_nodes_to_connect
from photon_b200.utils.core import get_n_cuda_devices
from photon_b200.utils.trace import tracer


def split_devices(n_devices: int, n_nodes: int) -> list[list[int] | None]:
    """Contiguous, near-equal device groups, one per node (``None`` = CPU node)."""
    if n_devices <= 0:
        return [None] * n_nodes
    if n_nodes > n_devices:
        raise ValueError(f"photon.n_nodes={n_nodes} exceeds the {n_devices} visible GPUs")
    per, extra = divmod(n_devices, n_nodes)
    out, lo = [], 0
    for i in range(n_nodes):
        hi = lo + per + (1 if i < extra else 0)
        out.append(list(range(lo, hi)))
        lo = hi
    return out


class NodeFleetRuntime(FederationRuntime):
    """Same interface as :class:`FederationRuntime` (``run_server`` drives either), different plumbing."""

    def __init__(self, cfg: Any, *, n_nodes: int | None = None, workers_per_node: int | None = None) -> None:
        super().__init__(cfg, device=torch.device("cpu"), rank=0, world_size=1)
        self.n_nodes = int(n_nodes if n_nodes is not None else cfg["photon"].get("n_nodes", 1))     # in-process nodes
        self.workers_per_node = workers_per_node
        self.apps: list[Any] = []            # in-process ClientApp objects and RemoteNode handles (same ``handle`` / ``alive`` interface)
        fleet = dict(cfg["photon"].get("fleet") or {})
        self.fleet_address = fleet.get("address") or None
        self.n_remote_nodes = int(fleet.get("n_remote_nodes", 0) or 0)
        self.fleet_liveness_s = float(fleet.get("liveness_timeout_s", 30.0) or 30.0)
        # hierarchical aggregation: nodes keep the weighted sum of the clients they trained, ONE model per node per round travels
        self.node_pre_aggregation = bool(fleet.get("node_pre_aggregation", False))
        self.link: Any = None
        self._pool: ThreadPoolExecutor | None = None
        self._uid = f"pb200_{uuid.uuid4().hex[:8]}"

    # --------------------------------------------------------------------- bring-up
    def build(self) -> None:
        _, self.model_layout = get_raw_model_parameters(self.cfg)      # CPU model → names/shapes only (ref: node_manager_app.py:261-271)
        self.layout = self.model_layout.stacked(("", "exp_avg/", "exp_avg_sq/")) if self.aggregate_momenta else self.model_layout
        self.round_backend = CollectiveRoundBackend(self.layout, self.strategy, self.device)   # world 1 → plain host server
        groups = split_devices(get_n_cuda_devices(), self.n_nodes) if self.n_nodes > 0 else []
        for i, devs in enumerate(groups):
            app = ClientApp(self.cfg, n_workers=self.workers_per_node or (len(devs) if devs else 1), node_id=i, devices=devs)
            app.nm.create_and_start_workers()
            self.apps.append(app)
        if self.n_remote_nodes > 0:
            import os

            fleet = dict(self.cfg["photon"].get("fleet") or {})
            from photon_b200.server.grpc_fleet import FleetLink

            tls = (os.environ.get("PHOTON_FLEET_TLS_KEY"), os.environ.get("PHOTON_FLEET_TLS_CERT"))
            self.link = FleetLink(self.fleet_address or "0.0.0.0:0", cfg=self.cfg, token=os.environ.get("PHOTON_FLEET_TOKEN"),
                                  liveness_timeout_s=self.fleet_liveness_s, tls=tls if all(tls) else None)
            print(f"[fleet] link listening on port {self.link.port}; waiting for {self.n_remote_nodes} remote node(s)", flush=True)
            self.apps += self.link.wait_for_nodes(self.n_remote_nodes, timeout_s=float(fleet.get("connect_timeout_s", 600.0) or 600.0))
        if not self.apps:
            raise ValueError("photon.topology=nodes needs photon.n_nodes >= 1 or photon.fleet.n_remote_nodes >= 1")
        self.n_nodes = len(self.apps)        # what the round loop waits for / reports: in-process + remote
        self._pool = ThreadPoolExecutor(max_workers=max(64, len(self.apps)), thread_name_prefix="node")   # room for nodes that join later
        wait_for_nodes_to_connect(len(self.apps), self.node_ids, poll_s=0.05, timeout_s=300.0)

    def node_ids(self) -> list[int]:
        self._admit_new_nodes()
        return [a.node_id for a in self.apps if a.alive()]

    def _admit_new_nodes(self) -> None:
        """Machines may join a running federation (ref: Flower's ``get_node_ids`` is whoever is connected NOW): a node that
        registered with the link since the last look gets a handle here; it receives the next broadcast and enters the work queue
        of the next round."""
        if self.link is None:
            return
        known = {a.node_id for a in self.apps}
        for node in self.link.nodes():
            if node.node_id not in known:
                self.apps.append(node)
                print(f"[fleet] node {node.node_id} joined the federation", flush=True)

    # --------------------------------------------------------------------- broadcast (R2)
    def broadcast_to_nodes(self) -> dict[str, Any]:
        """One payload on the side channel, one QUERY per node, wait for every ack."""
        assert self._pool is not None and self.round_backend is not None
        t0 = time.time()
        local = [a for a in self.apps if not getattr(a, "remote", False)]
        remote = [a for a in self.apps if getattr(a, "remote", False) and a.alive()]
        handles: list[ParamHandle] = []
        try:
            futs = []
            if local:       # one POSIX segment for every node of this machine
                h = replace_remote_with_parameters_in_recordset(
                    ParamHandle("inline", self.round_backend.global_params()), {"shm": True}, endpoint_id=f"{self._uid}_bcast", layout=self.layout)
                handles.append(h)
                futs += [self._pool.submit(app.handle, Message("query", {"type": "broadcast_parameters", "parameters": h}, node_id=app.node_id))
                         for app in local]
            if remote:      # one object for every other machine: in the S3 bucket, or in the link's own object service
                from photon_b200.utils.objstore import remote_store_from_cfg

                store = remote_store_from_cfg(self.cfg)
                inline = ParamHandle("inline", self.round_backend.global_params())
                h = replace_remote_with_parameters_in_recordset(
                    inline, "s3" if store is not None else "link", endpoint_id="server", layout=self.layout,
                    store=store if store is not None else self.link.spool, bucket=str(self.cfg["s3_comm_config"]["bucket_name"]),
                    folder_name=f"{self.cfg['run_uuid']}/server/comm_stack")
                handles.append(h)
                futs += [self._pool.submit(app.handle, Message("query", {"type": "broadcast_parameters", "parameters": h}, node_id=app.node_id))
                         for app in remote]
            targets = local + remote
            acks = [f.result() for f in futs]
        finally:
            for h in handles:
                release_remote_parameters(h)
        # what crossed the network this round (in-process nodes share a POSIX segment): the quantity the reference's design is about
        self.timings["comm/param_bytes_to_nodes"] = float(4 * self.layout.total * len(remote))
        self.timings["comm/param_bytes_from_nodes"] = 0.0
        # a remote node that died during the broadcast is simply out of the rotation (its liveness check fails from now on);
        # anything else that does not acknowledge is an error
        bad = [a.error or f"unexpected reply {a.content!r}" for app, a in zip(targets, acks)
               if (a.error or a.content != {"broadcast": {"status": "OK"}}) and not (getattr(app, "remote", False) and not app.alive())]
        if bad:
            raise RuntimeError(f"broadcast not acknowledged by {len(bad)} node(s): {bad[:2]}")
        if not any(a.content == {"broadcast": {"status": "OK"}} for a in acks):
            raise RuntimeError("no node of the fleet received the round's parameters")
        return {"server/broadcast_time": time.time() - t0}

    # ---------------------------------------------------------------------------- fit
    def run_clients_fit(self, server_round: int, sampled: list[int]) -> list[FitRes]:
        assert self._pool is not None and self.round_backend is not None
        rb = self.round_backend
        rb.begin_round()
        self.timings.update(self.broadcast_to_nodes())
        pending: dict[int, Future[Message]] = {}
        by_id = {a.node_id: a for a in self.apps}

        def dispatch(node: int, cid: int) -> None:
            if self._should_fail(server_round, cid):
                f: Future[Message] = Future()
                f.set_result(Message("train", [FitRes(Status(Code.FAILED, f"fault injection: client {cid} dropped in round {server_round}"),
                                                      None, 0, {}, cid)], node_id=node))
                pending[node] = f
                return
            # a node with capacity > 1 (a whole SPMD box: one client per GPU at a time) takes that many clients in ONE task
            cids = [cid] + [sched.queue.popleft() for _ in range(min(getattr(by_id[node], "capacity", 1) - 1, len(sched.queue)))]
            batches[node] = cids
            per = {c: self.fit_config_fn(server_round, c, self.client_states, self.server_steps_cumulative).to_wire() for c in cids}
            msg = fit_or_evaluate_ins("train", server_round, cids, self.client_states, self.server_steps_cumulative, per)
            msg.node_id = node
            msg.defer_parameters = self.node_pre_aggregation  # type: ignore[attr-defined]
            pending[node] = self._pool.submit(by_id[node].handle, msg)

        def poll() -> list[tuple[int, int, Message]]:
            done = [(n, f) for n, f in pending.items() if f.done()]
            out = []
            for n, f in done:
                del pending[n]
                rep = f.result()
                cid = rep.content[0].cid if rep.content else -1
                extra = batches.pop(n, [])[1:]
                if extra and getattr(by_id[n], "remote", False) and not by_id[n].alive():
                    sched.queue.extendleft(reversed(extra))      # the scheduler re-queues the first client of a lost node itself
                out.append((n, int(cid if cid is not None else -1), rep))
            return out

        results: list[FitRes] = []
        held: dict[int, list[int]] = {}       # node -> positions in ``results`` whose parameters the node is holding back
        batches: dict[int, list[int]] = {}    # node -> the clients of the task it is working on
        t0 = time.time()
        # a REMOTE node that stopped polling leaves the rotation (its clients are re-queued); an in-process node manager repairs its
        # own workers and reports what it could not do
        sched = ClientScheduler(sampled, [a.node_id for a in self.apps if a.alive()], dispatch, poll, poll_s=0.01,
                                is_alive=lambda n: by_id[n].alive() if getattr(by_id[n], "remote", False) else True)
        with tracer().span("fit_clients", cat="server", server_round=server_round):
            for _node, cid, reply in sched:
                for res in reply.content or [FitRes(Status(Code.FAILED, reply.error or "empty reply"), None, 0, {}, cid)]:
                    if res.status.code == Code.OK and res.parameters is not None and res.parameters.kind == "deferred":
                        held.setdefault(_node, []).append(len(results))
                        res = FitRes(res.status, ParamHandle(kind=rb.name), res.num_examples, res.metrics, res.cid)
                    elif res.status.code == Code.OK and res.parameters is not None:
                        if getattr(by_id[_node], "remote", False):
                            self.timings["comm/param_bytes_from_nodes"] += float(4 * self.layout.total)
                        flat = torch.zeros(self.layout.total, dtype=torch.float32)
                        if res.parameters.kind != "inline":     # a remote node parked its result in the bucket
                            from photon_b200.server.s3_utils import replace_parameters_in_recordset_with_remote

                            parked = res.parameters
                            res.parameters = replace_parameters_in_recordset_with_remote(parked)
                            release_remote_parameters(parked)
                        self.layout.from_ndarrays(flat, res.parameters.data)   # streaming: fold in, then drop the payload
                        rb.add_client(flat, res.num_examples)
                        res = FitRes(res.status, ParamHandle(kind=rb.name), res.num_examples, res.metrics, res.cid)
                    results.append(res)
        if held:
            self._collect_node_aggregates(server_round, held, results, by_id)
        self.timings["node_training_time_s"] = time.time() - t0
        return results

    def _collect_node_aggregates(self, server_round: int, held: dict[int, list[int]], results: list[FitRes], by_id: dict[int, Any]) -> None:
        """End of the round under ``node_pre_aggregation``: one ``collect_aggregate`` query per node that holds results; its weighted
        mean is folded in with the node's total weight. A node that cannot deliver (died after training) turns the clients it
        held into failures — the round's failure policy decides."""
        assert self._pool is not None and self.round_backend is not None
        from photon_b200.server.s3_utils import replace_parameters_in_recordset_with_remote

        futs = {n: self._pool.submit(by_id[n].handle, Message("query", {"type": "collect_aggregate", "server_round": server_round}, node_id=n))
                for n in held}
        for n, f in futs.items():
            rep = f.result()
            body = rep.content if isinstance(rep.content, dict) else {}
            want = sum(results[i].num_examples for i in held[n])
            if rep.error or body.get("aggregate") is None or int(body.get("num_examples", 0)) != want:
                why = rep.error or f"node {n} delivered {body.get('num_examples', 0)} of {want} examples"
                for i in held[n]:
                    results[i] = FitRes(Status(Code.FAILED, f"pre-aggregated result lost: {why}"), None, 0, {}, results[i].cid)
                continue
            handle = body["aggregate"]
            if handle.kind != "inline":
                parked = handle
                handle = replace_parameters_in_recordset_with_remote(parked)
                release_remote_parameters(parked)
            flat = torch.zeros(self.layout.total, dtype=torch.float32)
            self.layout.from_ndarrays(flat, handle.data)
            self.round_backend.add_client(flat, want)
            if getattr(by_id[n], "remote", False):
                self.timings["comm/param_bytes_from_nodes"] += float(4 * self.layout.total)

    # ----------------------------------------------------------------------- evaluate
    def run_clients_evaluate(self, server_round: int, sampled: list[int]) -> list[EvaluateRes]:
        self.broadcast_to_nodes()
        per = {int(cid): self.eval_config_fn(server_round, cid, self.client_states, self.server_steps_cumulative).to_wire() for cid in sampled}
        msg = fit_or_evaluate_ins("evaluate", server_round, list(per), self.client_states, self.server_steps_cumulative, per)
        first = next((a for a in self.apps if a.alive()), self.apps[0])
        rep = first.handle(msg)     # the reference evaluates on ONE node (client id 0, all streams concatenated)
        res = rep.content
        return [res] if isinstance(res, EvaluateRes) else [EvaluateRes(Status(Code.FAILED, rep.error or "no result"), 0.0, 0, {})]

    # --------------------------------------------------------------------------- close
    def close(self) -> None:
        for app in self.apps:
            app.shutdown()
        self.apps = []
        if self.link is not None:
            self.link.close()
            self.link = None
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        if self.round_backend is not None:
            self.round_backend.close()

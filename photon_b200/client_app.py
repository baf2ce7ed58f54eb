"""Node-side message handlers: ``train`` / ``evaluate`` / ``query`` / ``lifespan``
(ref: photon/client_app.py:78-291), bound to a :class:`NodeManagerApp`.

In the reference these are Flower ``ClientApp`` callbacks fed by gRPC; here they consume the
in-memory :class:`photon_b200.messages.Message` objects of our control plane and reply with the
same record fields (``fitres.*``, ``evaluateres.*``, ``{"broadcast": {"status": "OK"}}``).
"""
from __future__ import annotations

import contextlib
from typing import Any, Iterator

from photon_b200.messages import Code, EvaluateRes, FitRes, Message, ParamHandle, Status
from photon_b200.node_manager.node_manager_app import NodeManagerApp


class ClientApp:
    def __init__(self, cfg: Any, n_workers: int | None = None, node_id: int = 0, devices: list[int] | None = None) -> None:
        self.cfg, self.node_id = cfg, node_id
        self.nm = NodeManagerApp(cfg, n_workers=n_workers, devices=devices)
        self.refresh_period = int(cfg["photon"].get("refresh_period", 50))
        # node-level pre-aggregation (``photon.fleet.node_pre_aggregation``): Σ n_k·x_k and Σ n_k of the clients this node trained in
        # the current round; only the node's weighted mean travels to the server, once, when it asks (``collect_aggregate``)
        self._agg: list[Any] | None = None
        self._agg_w = 0
        self._agg_round = -1

    def alive(self) -> bool:
        """Every worker process of this node is up (what makes the node show up in ``node_ids()``)."""
        return bool(self.nm.workers) and all(w.is_alive() for w in self.nm.workers)

    def start(self) -> None:
        self.nm.create_and_start_workers()

    def describe(self) -> str:
        return f"{len(self.nm.workers)} worker(s)"

    def shutdown(self) -> None:
        self.nm.close()

    @contextlib.contextmanager
    def lifespan(self) -> Iterator["ClientApp"]:
        """Workers live as long as the app (the reference's ``--persist-client`` fork feature)."""
        self.nm.create_and_start_workers()
        try:
            yield self
        finally:
            self.nm.close()

    # ------------------------------------------------------------------------- handlers
    def set_parameters(self, msg: Message) -> Message:
        """Broadcast sink: fetch the payload from the side channel (shm / npz object) when the message carries a
        locator instead of the arrays, then publish it to this node's workers (ref: client_app.py:78-118)."""
        payload = msg.content["parameters"]
        if isinstance(payload, ParamHandle):
            from photon_b200.server.s3_utils import replace_parameters_in_recordset_with_remote

            payload = replace_parameters_in_recordset_with_remote(payload).data
        self.nm.set_parameters(payload)
        return Message("query", {"broadcast": {"status": "OK"}}, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)

    def train(self, msg: Message) -> Message:
        ins = msg.content
        server_round = int(ins.config["server_round"])
        if self.refresh_period and server_round > 1 and server_round % self.refresh_period == 0:
            self.nm.refresh_workers()                     # shed leaked memory (ref: client_app.py:175-177)
        per_client = getattr(msg, "per_client", {})
        results = self.nm.fit({cid: {"fit_config": fc} for cid, fc in per_client.items()})
        if getattr(msg, "defer_parameters", False):
            results = [self._fold(r, server_round) for r in results]
        return Message("train", results, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)

    def _fold(self, res: FitRes, server_round: int) -> FitRes:
        """Keep a successful client's parameters HERE (running weighted sum of the round); the reply carries metrics only."""
        import numpy as np

        if res.status.code != Code.OK or res.parameters is None:
            return res
        if self._agg_round != server_round:
            self._agg, self._agg_w, self._agg_round = None, 0, server_round
        n = int(res.num_examples)
        arrays = [np.asarray(a, dtype=np.float64) * n for a in res.parameters.data]
        self._agg = arrays if self._agg is None else [x + y for x, y in zip(self._agg, arrays)]
        self._agg_w += n
        return FitRes(res.status, ParamHandle("deferred", None, {"node_id": self.node_id}), res.num_examples, res.metrics, res.cid)

    def collect_aggregate(self, msg: Message) -> Message:
        """Hand over (weighted mean of this round's clients, Σ n_k) and forget it."""
        import numpy as np

        want = int((msg.content or {}).get("server_round", self._agg_round))
        if self._agg is None or self._agg_round != want or self._agg_w <= 0:
            body: Any = {"aggregate": None, "num_examples": 0}
        else:
            body = {"aggregate": ParamHandle("inline", [(a / self._agg_w).astype(np.float32) for a in self._agg]), "num_examples": self._agg_w}
        self._agg, self._agg_w = None, 0
        return Message("query", body, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)

    def evaluate(self, msg: Message) -> Message:
        per_client = getattr(msg, "per_client", {})
        res = self.nm.eval({cid: {"eval_config": ec} for cid, ec in per_client.items()})
        return Message("evaluate", res, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)

    def query(self, msg: Message) -> Message:
        kind = (msg.content or {}).get("type")
        if kind == "broadcast_parameters":
            return self.set_parameters(msg)
        if kind == "collect_aggregate":
            return self.collect_aggregate(msg)
        if kind == "free_resources":
            self.nm.refresh_workers()
            return Message("query", {"free_resources": {"status": "OK"}}, node_id=self.node_id, group_id=msg.group_id, reply_to=msg.msg_id)
        return Message("query", None, node_id=self.node_id, error=f"unknown query type {kind!r}", reply_to=msg.msg_id)

    def handle(self, msg: Message) -> Message:
        try:
            return {"train": self.train, "evaluate": self.evaluate, "query": self.query}[msg.kind](msg)
        except Exception as e:  # noqa: BLE001 - a node never crashes the federation; it replies with an error
            bad: Any = [FitRes(Status(Code.FAILED, repr(e)), None, 0, {})] if msg.kind == "train" else (
                EvaluateRes(Status(Code.FAILED, repr(e)), 0.0, 0, {}) if msg.kind == "evaluate" else None)
            return Message(msg.kind, bad, node_id=self.node_id, error=repr(e), reply_to=msg.msg_id)


# ----------------------------------------------------------------------------- module-level callbacks (ref: client_app.py:78-291)
# The reference registers plain functions on a Flower ``ClientApp`` (``@app.train()`` …) that share a module-global node manager
# created by ``lifespan``. Same shape: ``lifespan(cfg)`` builds THE app of this process, the functions below dispatch to it.
_APP: ClientApp | None = None


def _app() -> ClientApp:
    if _APP is None:
        raise RuntimeError("no node app in this process: wrap the calls in `with lifespan(cfg):` (or build a ClientApp yourself)")
    return _APP


@contextlib.contextmanager
def lifespan(cfg: Any = None, n_workers: int | None = None, devices: list[int] | None = None, node_id: int = 0) -> Iterator[ClientApp]:
    """Workers of this process's node live inside the ``with`` block. ``cfg`` defaults to ``$PHOTON_SAVE_PATH/config.yaml``."""
    global _APP
    if cfg is None:
        import os
        from pathlib import Path

        from photon_b200.config import load_config

        cfg = load_config(Path(os.environ["PHOTON_SAVE_PATH"]) / "config.yaml")
    app = ClientApp(cfg, n_workers=n_workers, node_id=node_id, devices=devices)
    with app.lifespan():
        _APP = app
        try:
            yield app
        finally:
            _APP = None


def train(msg: Message, context: Any = None) -> Message:
    del context
    return _app().handle(msg) if msg.kind == "train" else _app().train(msg)


def evaluate(msg: Message, context: Any = None) -> Message:
    del context
    return _app().evaluate(msg)


def query(msg: Message, context: Any = None) -> Message:
    del context
    return _app().query(msg)


def set_parameters(msg: Message, context: Any = None) -> Message:
    del context
    return _app().set_parameters(msg)


def free_resources(msg: Message | None = None, context: Any = None) -> Message:
    """Recycle the node's worker processes (ref: client_app.py ``free_resources`` query)."""
    del context
    return _app().query(Message("query", {"type": "free_resources"}, node_id=_app().node_id, reply_to=getattr(msg, "msg_id", None)))

"""Cross-host node fleet over gRPC: the reference's SuperLink ↔ SuperNode link.

The reference deploys one Flower SuperLink next to the server app and one SuperNode per machine; nodes CONNECT to the link and
pull their work from it, so only the server needs a reachable address (ref: scripts/photon_llm.sh launch order, photon/client_app.py,
photon/server_app.py:285,346 — ``driver.get_node_ids()`` is whoever is connected right now). Same shape here, on ``grpcio`` with
generic (bytes-in / bytes-out) method handlers — no generated stubs:

* :class:`FleetLink` (server process): ``Register`` hands a node its id AND the run's composed config (a node needs nothing but
  the server's address), ``Pull`` is a long poll for the node's next :class:`~photon_b200.messages.Message`, ``Push`` returns the
  reply, ``Beat`` is the heartbeat of a node that is busy training. Every call refreshes the node's liveness; a node silent for ``liveness_timeout_s`` drops out of ``node_ids()`` and whatever
  it was running comes back as a FAILED result, which the round loop already knows how to handle (``accept_failures_cnt``,
  re-queue on another node).
* :class:`RemoteNode`: the server-side stand-in with the ``handle(msg) -> Message`` interface of an in-process
  :class:`~photon_b200.client_app.ClientApp` — :class:`~photon_b200.server.fleet.NodeFleetRuntime` schedules local and remote
  nodes through the same work queue.
* :func:`serve_node` (node process, ``python -m photon_b200.node``): registers, builds its ``ClientApp`` (node manager + one
  worker per local GPU, DDP / ZeRO inside the node), then pull → handle → push until the server says ``shutdown``.

Parameters never ride inside a control message: with an object store configured (``S3_ENDPOINT_URL`` …) both directions park them in
the bucket and send the key (``ParamHandle("s3", key)``), like the reference's ``comm_stack.s3``; without one the LINK ITSELF is the
object store — the role Ray's object store plays in the reference's ``comm_stack.ray``: the server spools npz objects in a scratch
directory, nodes pull / push them in 32 MiB chunks over three more RPCs (``ObjGet`` / ``ObjPut`` / ``ObjDel``,
:class:`LinkObjectStore`), the message carries ``ParamHandle("link", key)``. No size limit (a gRPC frame is never larger than a chunk).

Messages are pickled: the link is for a trusted cluster (as the reference's pickled payloads are). ``PHOTON_FLEET_TOKEN`` — when set
on both sides — is checked on every call; TLS: pass ``tls=(key, cert)`` / ``tls_ca`` (env ``PHOTON_FLEET_TLS_KEY`` / ``_CERT`` / ``_CA``).
"""
from __future__ import annotations

import os
import pickle
import queue
import threading
import time
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any

from photon_b200.messages import Code, EvaluateRes, FitRes, Message, ParamHandle, Status

SERVICE = "photon.Fleet"
_MAX_MSG = 2**31 - 1
_OPTS = [("grpc.max_send_message_length", _MAX_MSG), ("grpc.max_receive_message_length", _MAX_MSG),
         ("grpc.keepalive_time_ms", 30000), ("grpc.keepalive_permit_without_calls", 1)]


def _ser(obj: Any) -> bytes:
    return pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)


def _de(data: bytes) -> Any:
    return pickle.loads(data)  # noqa: S301 - trusted cluster link (see module docstring)


def _failed(msg: Message, why: str) -> Message:
    bad: Any = None
    if msg.kind == "train":
        cids = list(getattr(msg, "per_client", {}) or [None])
        bad = [FitRes(Status(Code.FAILED, why), None, 0, {}, c) for c in cids]
    elif msg.kind == "evaluate":
        bad = EvaluateRes(Status(Code.FAILED, why), 0.0, 0, {})
    return Message(msg.kind, bad, node_id=msg.node_id, error=why, reply_to=msg.msg_id)


class _Slot:
    def __init__(self, node_id: int, info: dict[str, Any]) -> None:
        self.node_id, self.info = node_id, info
        self.outbox: "queue.Queue[Message]" = queue.Queue()
        self.pending: dict[int, tuple[Message, Future]] = {}
        self.last_seen = time.time()
        self.lock = threading.Lock()
        self.gone = False


class RemoteNode:
    """Server-side handle of a connected node (``ClientApp`` interface)."""

    def __init__(self, link: "FleetLink", slot: _Slot) -> None:
        self._link, self._slot = link, slot
        self.node_id = slot.node_id
        self.remote = True
        self.capacity = max(1, int(slot.info.get("capacity", 1) or 1))     # clients the node takes per task (a whole SPMD box: its GPUs)

    def alive(self) -> bool:
        return not self._slot.gone and time.time() - self._slot.last_seen < self._link.liveness_timeout_s

    def handle(self, msg: Message) -> Message:
        """Send ``msg`` to the node and wait for its reply; a node that stops polling turns into a FAILED reply."""
        s = self._slot
        fut: Future = Future()
        with s.lock:
            msg.msg_id = self._link.next_msg_id()
            msg.node_id = self.node_id
            s.pending[msg.msg_id] = (msg, fut)
        s.outbox.put(msg)
        while True:
            try:
                return fut.result(timeout=1.0)
            except TimeoutError:
                if not self.alive():
                    with s.lock:
                        s.pending.pop(msg.msg_id, None)
                    return _failed(msg, f"node {self.node_id} stopped responding ({time.time() - s.last_seen:.0f} s since its last call)")

    def shutdown(self) -> None:
        if self.alive():
            self._slot.outbox.put(Message("shutdown", None, node_id=self.node_id))


class FleetLink:
    def __init__(self, address: str = "0.0.0.0:0", *, cfg: Any = None, token: str | None = None, liveness_timeout_s: float = 30.0,
                 tls: tuple[str, str] | None = None) -> None:
        import grpc

        self.cfg, self.token, self.liveness_timeout_s = cfg, token, float(liveness_timeout_s)
        self._slots: dict[int, _Slot] = {}
        self._lock = threading.Lock()
        self._msg_id = 0
        import tempfile

        from photon_b200.server import s3_utils
        from photon_b200.utils.objstore import DirObjectStore

        self._spool_dir = tempfile.mkdtemp(prefix="photon_link_")
        self.spool = DirObjectStore(self._spool_dir)           # what ObjGet / ObjPut / ObjDel serve; the server reads it directly
        s3_utils.set_link_store(self.spool)
        # every connected node parks one long poll here (plus its heartbeats and object chunks): threads, not work
        self._server = grpc.server(ThreadPoolExecutor(max_workers=int(os.environ.get("PHOTON_FLEET_THREADS", "256")), thread_name_prefix="fleet-link"),
                                   options=_OPTS)
        ident = lambda b: b  # noqa: E731 - payloads are already bytes
        handlers = {name: grpc.unary_unary_rpc_method_handler(getattr(self, "_" + name.lower()), request_deserializer=ident, response_serializer=ident)
                    for name in ("Register", "Pull", "Push", "Beat", "ObjGet", "ObjPut", "ObjDel")}
        self._server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE, handlers),))
        if tls is not None:
            key, cert = (open(p, "rb").read() for p in tls)
            self.port = self._server.add_secure_port(address, grpc.ssl_server_credentials([(key, cert)]))
        else:
            self.port = self._server.add_insecure_port(address)
        if not self.port:
            raise RuntimeError(f"fleet link could not bind {address}")
        host = address.rsplit(":", 1)[0]
        if not token and host not in ("127.0.0.1", "localhost", "[::1]"):
            # messages are pickles: whoever can reach the port can run code in this process. Loopback needs nothing; anything else
            # should carry the shared token (checked before a single byte is unpickled) and, across sites, TLS.
            print(f"[fleet] WARNING: the link listens on {address} without PHOTON_FLEET_TOKEN — any host that can reach it is trusted",
                  flush=True)
        self._server.start()

    # ------------------------------------------------------------------ rpc handlers
    def _auth(self, context: Any) -> None:
        import grpc

        if self.token and dict(context.invocation_metadata()).get("x-photon-token") != self.token:
            context.abort(grpc.StatusCode.UNAUTHENTICATED, "bad fleet token")

    def _slot(self, node_id: int, context: Any) -> _Slot:
        import grpc

        s = self._slots.get(int(node_id))
        if s is None or s.gone:
            context.abort(grpc.StatusCode.NOT_FOUND, f"node {node_id} is not registered")
        s.last_seen = time.time()
        return s

    def _register(self, request: bytes, context: Any) -> bytes:
        self._auth(context)
        info = _de(request)
        with self._lock:
            node_id = max(list(self._slots) + [int(info.get("first_id", 1000)) - 1]) + 1
            self._slots[node_id] = _Slot(node_id, info)
        print(f"[fleet] node {node_id} registered: {info.get('host')} devices={info.get('devices')} workers={info.get('n_workers')}", flush=True)
        return _ser({"node_id": node_id, "cfg": self.cfg, "liveness_timeout_s": self.liveness_timeout_s})

    def _beat(self, request: bytes, context: Any) -> bytes:
        """Heartbeat of a node that is busy training (it does not poll while it works)."""
        self._auth(context)
        self._slot(_de(request)["node_id"], context)
        return _ser(True)

    def _pull(self, request: bytes, context: Any) -> bytes:
        self._auth(context)
        req = _de(request)
        s = self._slot(req["node_id"], context)
        try:
            msg = s.outbox.get(timeout=float(req.get("wait_s", 2.0)))
        except queue.Empty:
            return _ser(None)
        s.last_seen = time.time()
        return _ser(msg)

    def _push(self, request: bytes, context: Any) -> bytes:
        self._auth(context)
        req = _de(request)
        s = self._slot(req["node_id"], context)
        reply: Message = req["reply"]
        with s.lock:
            entry = s.pending.pop(int(reply.reply_to) if reply.reply_to is not None else -1, None)
        if entry is not None:
            entry[1].set_result(reply)
        return _ser(True)

    # ------------------------------------------------------------------ object service (the link as the parameter store)
    def _obj_path(self, key: str, context: Any) -> Any:
        import grpc
        from pathlib import Path

        p = (Path(self._spool_dir) / key).resolve()
        if Path(self._spool_dir).resolve() not in p.parents:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, "key escapes the spool")
        return p

    def _objget(self, request: bytes, context: Any) -> bytes:
        import grpc

        self._auth(context)
        req = _de(request)
        p = self._obj_path(req["key"], context)
        if not p.is_file():
            context.abort(grpc.StatusCode.NOT_FOUND, f"no such object: {req['key']}")
        with open(p, "rb") as f:
            f.seek(int(req["offset"]))
            data = f.read(min(int(req["n"]), 256 << 20))      # a frame stays far below gRPC's 2 GiB limit whatever the caller asks
        return _ser({"data": data, "size": p.stat().st_size})

    def _objput(self, request: bytes, context: Any) -> bytes:
        self._auth(context)
        req = _de(request)
        p = self._obj_path(req["key"], context)
        part = p.with_name(p.name + ".part")
        p.parent.mkdir(parents=True, exist_ok=True)
        with open(part, "r+b" if int(req["offset"]) and part.exists() else "wb") as f:
            f.seek(int(req["offset"]))
            f.write(req["data"])
        if req.get("last"):
            os.replace(part, p)
        return _ser(True)

    def _objdel(self, request: bytes, context: Any) -> bytes:
        self._auth(context)
        self.spool.delete(_de(request)["key"])
        return _ser(True)

    # ------------------------------------------------------------------ server API
    def next_msg_id(self) -> int:
        with self._lock:
            self._msg_id += 1
            return self._msg_id

    def nodes(self) -> list[RemoteNode]:
        return [RemoteNode(self, s) for s in self._slots.values() if not s.gone]

    def wait_for_nodes(self, n: int, timeout_s: float = 600.0, poll_s: float = 0.1) -> list[RemoteNode]:
        """Block until ``n`` nodes have registered (ref: photon/server/server_util.py ``wait_for_nodes_to_connect``)."""
        t0 = time.time()
        while len(self._slots) < n:
            if time.time() - t0 > timeout_s:
                raise TimeoutError(f"{len(self._slots)} of {n} remote nodes connected to the fleet link within {timeout_s:.0f} s")
            time.sleep(poll_s)
        return self.nodes()

    def close(self, grace_s: float = 2.0) -> None:
        for n in self.nodes():
            n.shutdown()
        t0 = time.time()
        while time.time() - t0 < grace_s and any(not s.outbox.empty() for s in self._slots.values()):
            time.sleep(0.05)
        for s in self._slots.values():
            s.gone = True
        self._server.stop(grace=grace_s)
        import shutil

        from photon_b200.server import s3_utils

        if s3_utils.get_link_store() is self.spool:
            s3_utils.set_link_store(None)
        shutil.rmtree(self._spool_dir, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------------------- node side
class LinkObjectStore:
    """The fleet link's object service seen from a node (same interface as :mod:`photon_b200.utils.objstore`)."""

    CHUNK = int(os.environ.get("PHOTON_LINK_CHUNK", 32 << 20))

    def __init__(self, call: dict[str, Any], md: Any) -> None:
        self._call, self._md = call, md

    def _rpc(self, name: str, body: dict[str, Any]) -> Any:
        return _de(self._call[name](_ser(body), metadata=self._md, timeout=3600.0))

    def upload(self, key: str, path: Any) -> None:
        size, off = os.path.getsize(path), 0
        with open(path, "rb") as f:
            while True:
                data = f.read(self.CHUNK)
                last = off + len(data) >= size
                self._rpc("ObjPut", {"key": key, "offset": off, "data": data, "last": last})
                off += len(data)
                if last:
                    return

    def put(self, key: str, data: bytes) -> None:
        for off in range(0, max(len(data), 1), self.CHUNK):
            self._rpc("ObjPut", {"key": key, "offset": off, "data": data[off: off + self.CHUNK], "last": off + self.CHUNK >= len(data)})

    def download(self, key: str, path: Any) -> Any:
        from pathlib import Path

        p = Path(path)
        p.parent.mkdir(parents=True, exist_ok=True)
        tmp = p.with_name(f"{p.name}.{os.getpid()}.part")
        off = 0
        with open(tmp, "wb") as f:
            while True:
                r = self._rpc("ObjGet", {"key": key, "offset": off, "n": self.CHUNK})
                f.write(r["data"])
                off += len(r["data"])
                if off >= r["size"] or not r["data"]:
                    break
        os.replace(tmp, p)
        return p

    def get(self, key: str) -> bytes:
        out, off = bytearray(), 0
        while True:
            r = self._rpc("ObjGet", {"key": key, "offset": off, "n": self.CHUNK})
            out += r["data"]
            off += len(r["data"])
            if off >= r["size"] or not r["data"]:
                return bytes(out)

    def delete(self, key: str) -> None:
        self._rpc("ObjDel", {"key": key})


def _park_results(reply: Message, cfg: Any, node_id: int) -> Message:
    """Node → server parameters through the object store when there is one (key in the message instead of the arrays)."""
    from photon_b200.utils.objstore import remote_store_from_cfg

    from photon_b200.server import s3_utils

    store, kind, meta = remote_store_from_cfg(cfg), "s3", {"bucket": str(cfg["s3_comm_config"]["bucket_name"])}
    if store is None:
        store, kind, meta = s3_utils.get_link_store(), "link", {}
    if store is None:
        return reply
    import tempfile
    from pathlib import Path

    from photon_b200.utils.core import dump_model_parameters_to_file

    if reply.kind == "query" and isinstance(reply.content, dict) and isinstance(reply.content.get("aggregate"), ParamHandle) \
            and reply.content["aggregate"].kind == "inline":       # the node's pre-aggregated round result
        key = f"{cfg['run_uuid']}/server/comm_stack/node-{node_id}/aggregate.npz"
        with tempfile.TemporaryDirectory() as td:
            store.upload(key, dump_model_parameters_to_file(Path(td) / "p.npz", list(reply.content["aggregate"].data)))
        reply.content["aggregate"] = ParamHandle(kind, key, dict(meta))
        return reply
    if reply.kind != "train" or not isinstance(reply.content, list):
        return reply

    for res in reply.content:
        if isinstance(res, FitRes) and res.parameters is not None and res.parameters.kind == "inline":
            key = f"{cfg['run_uuid']}/server/comm_stack/node-{node_id}/client_{res.cid}.npz"
            with tempfile.TemporaryDirectory() as td:
                p = dump_model_parameters_to_file(Path(td) / "p.npz", list(res.parameters.data))
                store.upload(key, p)
            res.parameters = ParamHandle(kind, key, dict(meta))
    return reply


def serve_node(server_address: str, *, n_workers: int | None = None, devices: list[int] | None = None, token: str | None = None,
               tls_ca: str | None = None, first_id: int = 1000, max_idle_s: float | None = None, app_factory: Any = None,
               info: dict[str, Any] | None = None) -> int:
    """Run one node until the server shuts the fleet down (or is unreachable for ``max_idle_s``). Returns the node id.
    ``app_factory(cfg, node_id)`` builds what serves the messages (default: a :class:`ClientApp` — node manager + workers;
    :mod:`photon_b200.server.box_node` passes the SPMD runtime of a whole box); ``info`` is added to the registration record
    (``capacity``: clients the node takes per task)."""
    import socket

    import grpc

    token = token if token is not None else os.environ.get("PHOTON_FLEET_TOKEN")
    tls_ca = tls_ca or os.environ.get("PHOTON_FLEET_TLS_CA")
    channel = (grpc.secure_channel(server_address, grpc.ssl_channel_credentials(open(tls_ca, "rb").read()), options=_OPTS) if tls_ca
               else grpc.insecure_channel(server_address, options=_OPTS))
    md = (("x-photon-token", token),) if token else ()
    ident = lambda b: b  # noqa: E731
    call = {name: channel.unary_unary(f"/{SERVICE}/{name}", request_serializer=ident, response_deserializer=ident)
            for name in ("Register", "Pull", "Push", "Beat", "ObjGet", "ObjPut", "ObjDel")}
    from photon_b200.server import s3_utils

    s3_utils.set_link_store(LinkObjectStore(call, md))       # "link" parameter handles resolve through this node's connection
    t0 = time.time()
    while True:     # the server may come up after its nodes (the reference starts the supernodes in any order)
        try:
            reg = _de(call["Register"](_ser({"host": socket.gethostname(), "pid": os.getpid(), "n_workers": n_workers, "devices": devices, "first_id": first_id, **(info or {})}),
                                       metadata=md, timeout=30.0))
            break
        except grpc.RpcError as e:
            if e.code() == grpc.StatusCode.UNAUTHENTICATED or time.time() - t0 > (max_idle_s or 600.0):
                raise
            time.sleep(1.0)
    node_id, cfg = int(reg["node_id"]), reg["cfg"]
    if app_factory is not None:
        app = app_factory(cfg, node_id)
    else:
        from photon_b200.client_app import ClientApp

        app = ClientApp(cfg, n_workers=n_workers, node_id=node_id, devices=devices)
    app.start()
    print(f"[node {node_id}] connected to {server_address}; {app.describe()}", flush=True)
    last_ok = time.time()
    stop = threading.Event()

    def heartbeat() -> None:     # a training task takes minutes; the link must keep hearing from this node meanwhile
        period = max(0.2, float(reg.get("liveness_timeout_s", 30.0)) / 4.0)
        while not stop.wait(period):
            try:
                call["Beat"](_ser({"node_id": app.node_id}), metadata=md, timeout=10.0)
            except grpc.RpcError:
                pass

    hb = threading.Thread(target=heartbeat, name="photon-node-heartbeat", daemon=True)
    hb.start()
    try:
        while True:
            try:
                msg = _de(call["Pull"](_ser({"node_id": node_id, "wait_s": 2.0}), metadata=md, timeout=30.0))
                last_ok = time.time()
            except grpc.RpcError as e:
                if e.code() == grpc.StatusCode.NOT_FOUND:
                    # the link answers but does not know this node: the server process was restarted (crash + resume). Join the new
                    # instance with the workers / runtime that are already up, as long as it is the same run.
                    try:
                        reg = _de(call["Register"](_ser({"host": socket.gethostname(), "pid": os.getpid(), "n_workers": n_workers, "devices": devices,
                                                         "first_id": first_id, "rejoined_from": node_id, **(info or {})}), metadata=md, timeout=30.0))
                    except grpc.RpcError:
                        time.sleep(0.5)
                        continue
                    if str((reg["cfg"] or {}).get("run_uuid")) != str((cfg or {}).get("run_uuid")):
                        print(f"[node {node_id}] the link now serves another run ({(reg['cfg'] or {}).get('run_uuid')}); leaving", flush=True)
                        break
                    print(f"[node {node_id}] the server was restarted: re-registered as node {reg['node_id']}", flush=True)
                    node_id = int(reg["node_id"])
                    app.node_id = node_id
                    last_ok = time.time()
                    continue
                if max_idle_s is not None and time.time() - last_ok > max_idle_s:
                    print(f"[node {node_id}] lost the fleet link ({e.code().name}); leaving", flush=True)
                    break
                time.sleep(0.5)
                continue
            getattr(app, "tick", lambda: None)()
            if msg is None:
                continue
            if msg.kind == "shutdown":
                break
            reply = app.handle(msg)
            reply.reply_to = msg.msg_id
            reply = _park_results(reply, cfg, node_id)
            for attempt in range(5):
                try:
                    call["Push"](_ser({"node_id": node_id, "reply": reply}), metadata=md, timeout=3600.0)
                    break
                except grpc.RpcError:
                    if attempt == 4:
                        raise
                    time.sleep(1.0 + attempt)
    finally:
        stop.set()
        hb.join(timeout=15.0)       # never tear the channel down under a heartbeat that is inside an RPC (the C++ runtime aborts the process)
        app.shutdown()
        channel.close()
    return node_id

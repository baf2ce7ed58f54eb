"""Cross-host deployment shape: the server's gRPC fleet link + ``python -m photon_b200.node`` processes (here on 127.0.0.1).
Parameters inline in the gRPC messages, or through an S3 bucket; a node that dies mid-run leaves the rotation."""
import os
import socket
import subprocess
import sys
import threading
import time
from pathlib import Path

import pytest
import torch

from fake_s3 import ACCESS, REGION, SECRET, FakeS3
from test_federation_cpu import _cfg

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn_nodes(n, port, env=None, extra=()):
    e = dict(os.environ, PYTHONPATH=str(ROOT), CUDA_VISIBLE_DEVICES="", **(env or {}))
    return [subprocess.Popen([sys.executable, "-m", "photon_b200.node", "--server", f"127.0.0.1:{port}", "--n-workers", "1", "--max-idle-s", "60", *extra],
                             env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for _ in range(n)]


def _reap(procs):
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=60)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append(p.communicate()[0])
    return outs


COMMON = ["fl.n_rounds=2", "fl.n_clients_per_round=3", "llm_config.save_folder=null", "fl.strategy_name=fedavg", "photon.topology=nodes"]


_REF: dict[tuple, torch.Tensor] = {}      # the same in-process two-node run serves every variant of a scenario


def _reference_model(tmp_path, run_uuid="fleet", common=None):
    from photon_b200.server.fleet import NodeFleetRuntime
    from photon_b200.server_app import run_server

    key = (run_uuid, tuple(common or COMMON))
    if key not in _REF:
        cfg = _cfg(tmp_path / "local", f"run_uuid={run_uuid}", "photon.n_nodes=2", *(common or COMMON))
        rt = NodeFleetRuntime(cfg)
        try:
            run_server(cfg, runtime=rt)
            _REF[key] = rt.round_backend.global_params().clone()
        finally:
            rt.close()
    return _REF[key]


@pytest.mark.parametrize("store", ["link", "s3"])
def test_remote_nodes_over_grpc_match_local_nodes(tmp_path, store, monkeypatch):
    from photon_b200.server.fleet import NodeFleetRuntime
    from photon_b200.server_app import run_server

    want = _reference_model(tmp_path)
    s3 = FakeS3() if store == "s3" else None
    env = {"PHOTON_FLEET_TOKEN": "sesame"}
    if s3 is not None:
        env.update({"S3_ENDPOINT_URL": s3.endpoint, "AWS_ACCESS_KEY_ID": ACCESS, "AWS_SECRET_ACCESS_KEY": SECRET, "AWS_DEFAULT_REGION": REGION})
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    port = _free_port()
    if store == "link":
        env["PHOTON_LINK_CHUNK"] = "65536"      # many chunks per object: exercises the chunked ObjGet / ObjPut paths
        monkeypatch.setenv("PHOTON_LINK_CHUNK", "65536")
    procs = _spawn_nodes(2, port, env)          # nodes may start before the server: they keep trying to register
    cfg = _cfg(tmp_path / "remote", "run_uuid=fleet", "photon.n_nodes=0", f"photon.fleet.address=127.0.0.1:{port}", "photon.fleet.n_remote_nodes=2",
               "s3_comm_config.bucket_name=fleetbkt", *COMMON)
    rt = NodeFleetRuntime(cfg)
    calls = {"get": 0, "put": 0}
    from photon_b200.server import grpc_fleet

    og, op = grpc_fleet.FleetLink._objget, grpc_fleet.FleetLink._objput
    monkeypatch.setattr(grpc_fleet.FleetLink, "_objget", lambda self, r, c: (calls.__setitem__("get", calls["get"] + 1), og(self, r, c))[1])
    monkeypatch.setattr(grpc_fleet.FleetLink, "_objput", lambda self, r, c: (calls.__setitem__("put", calls["put"] + 1), op(self, r, c))[1])
    try:
        h = run_server(cfg, runtime=rt)
        fit = h.metrics_distributed_fit
        assert [v for _, v in fit["server/n_failures"]] == [0, 0] and [v for _, v in fit["server/n_nodes"]] == [2, 2]
        assert sorted(rt.node_ids()) == [1000, 1001] and len(h.losses_distributed) == 3
        got = rt.round_backend.global_params().clone()
    finally:
        rt.close()
    outs = _reap(procs)
    assert all(p.returncode == 0 for p in procs), outs
    assert torch.allclose(got, want, atol=1e-5), float((got - want).abs().max())     # replies arrive in any order: another fp32 summation order
    if store == "link":
        assert calls["get"] > 8 and calls["put"] > 8, calls      # broadcast pulled and results pushed through the link, chunk by chunk
    if s3 is not None:
        keys = [k for m, k in s3.requests if m == "PUT"]
        assert any("fleetbkt/fleet/server/comm_stack/server/parameters.npz" in k for k in keys)          # broadcast went through the bucket
        assert any("comm_stack/node-1000/client_" in k or "comm_stack/node-1001/client_" in k for k in keys)   # and so did the results
        assert not [k for k in s3.objects if "/comm_stack/" in k]                                       # receivers released what they fetched
        s3.stop()


@pytest.mark.parametrize("store", ["link", "s3"])
def test_node_pre_aggregation_ships_one_model_per_node(tmp_path, store, monkeypatch):
    """``photon.fleet.node_pre_aggregation``: four sampled clients on two remote nodes -> the train replies carry no parameters, each
    node answers ONE ``collect_aggregate`` query per round with its weighted mean, and the global model equals the per-client path."""
    from photon_b200.server.fleet import NodeFleetRuntime
    from photon_b200.server_app import run_server

    common = ["fl.n_rounds=2", "fl.n_clients_per_round=4", "llm_config.save_folder=null", "fl.strategy_name=fedavg", "photon.topology=nodes",
              "fl.eval_period=null"]
    want = _reference_model(tmp_path, "agg", common)
    s3 = FakeS3() if store == "s3" else None
    env = {}
    if s3 is not None:
        env = {"S3_ENDPOINT_URL": s3.endpoint, "AWS_ACCESS_KEY_ID": ACCESS, "AWS_SECRET_ACCESS_KEY": SECRET, "AWS_DEFAULT_REGION": REGION}
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    port = _free_port()
    procs = _spawn_nodes(2, port, env)
    cfg = _cfg(tmp_path / "remote", "run_uuid=agg", "photon.n_nodes=0", f"photon.fleet.address=127.0.0.1:{port}", "photon.fleet.n_remote_nodes=2",
               "photon.fleet.node_pre_aggregation=true", "s3_comm_config.bucket_name=aggbkt", *common)
    rt = NodeFleetRuntime(cfg)
    seen = {"train_with_params": 0, "collect": 0}
    try:
        from photon_b200.server import grpc_fleet

        orig = grpc_fleet.FleetLink._push

        def spy(self, request, context):
            reply = grpc_fleet._de(request)["reply"]
            if reply.kind == "train":
                seen["train_with_params"] += sum(1 for r in reply.content if r.parameters is not None and r.parameters.kind != "deferred")
            if reply.kind == "query" and isinstance(reply.content, dict) and "aggregate" in reply.content:
                seen["collect"] += 1
            return orig(self, request, context)

        monkeypatch.setattr(grpc_fleet.FleetLink, "_push", spy)
        h = run_server(cfg, runtime=rt)
        assert [v for _, v in h.metrics_distributed_fit["server/n_failures"]] == [0, 0]
        got = rt.round_backend.global_params().clone()
        size = 4 * rt.layout.total
        inbound = [v for _, v in h.metrics_centralized["comm/param_bytes_from_nodes"]]
        assert all(size <= b <= 2 * size for b in inbound), (inbound, size)            # at most one model per node, not one per client (4)
        assert [v for _, v in h.metrics_centralized["comm/param_bytes_to_nodes"]] == [2 * size, 2 * size]
    finally:
        rt.close()
    outs = _reap(procs)
    assert all(p.returncode == 0 for p in procs), outs
    assert seen["train_with_params"] == 0 and 2 <= seen["collect"] <= 4, seen       # <= one aggregate per node per round, never a client model
    assert torch.allclose(got, want, atol=1e-5), float((got - want).abs().max())     # a mean of node means: same value, another rounding order
    if s3 is not None:
        puts = [k for m, k in s3.requests if m == "PUT"]
        assert any("/node-1000/aggregate.npz" in k or "/node-1001/aggregate.npz" in k for k in puts) and not any("/client_" in k for k in puts)
        s3.stop()


def test_wrong_token_is_refused(tmp_path, monkeypatch):
    from photon_b200.server.grpc_fleet import FleetLink

    link = FleetLink("127.0.0.1:0", cfg={"x": 1}, token="right")
    try:
        procs = _spawn_nodes(1, link.port, {"PHOTON_FLEET_TOKEN": "wrong"})
        out = _reap(procs)[0]
        assert procs[0].returncode != 0 and "UNAUTHENTICATED" in out and not link.nodes()
    finally:
        link.close(grace_s=0.1)


def test_node_lost_mid_round_leaves_the_rotation(tmp_path, monkeypatch):
    """One of two remote nodes is killed while it trains: its client is re-queued on the survivor, the round completes with every
    sampled client, the dead node is gone from ``node_ids()`` afterwards."""
    from photon_b200.server.fleet import NodeFleetRuntime
    from photon_b200.server_app import run_server

    port = _free_port()
    procs = _spawn_nodes(2, port)
    cfg = _cfg(tmp_path, "run_uuid=lost", "photon.n_nodes=0", f"photon.fleet.address=127.0.0.1:{port}", "photon.fleet.n_remote_nodes=2",
               "photon.fleet.liveness_timeout_s=4", "photon.fleet.connect_timeout_s=120", "fl.n_rounds=2", "fl.n_clients_per_round=4", "llm_config.save_folder=null", "fl.strategy_name=fedavg",
               "photon.topology=nodes", "fl.eval_period=null", "llm_config.local_steps=6ba")
    rt = NodeFleetRuntime(cfg)

    def killer():
        t0 = time.time()
        while time.time() - t0 < 120:      # wait until some node holds a training task of round 1, then kill THAT node's process
            link = rt.link
            slots = list(link._slots.values()) if link is not None else []
            busy = [s for s in slots if any(m.kind == "train" for m, _ in list(s.pending.values()))]
            if len(slots) == 2 and busy:
                time.sleep(0.3)
                victim = busy[-1].info["pid"]
                next(p for p in procs if p.pid == victim).kill()
                return
            time.sleep(0.02)

    try:
        threading.Thread(target=killer, daemon=True).start()
        h = run_server(cfg, runtime=rt)
        fit = h.metrics_distributed_fit
        assert [v for _, v in fit["server/n_failures"]] == [0, 0]
        assert len(rt.node_ids()) == 1 and [v for _, v in fit["server/n_nodes"]][-1] == 1
    finally:
        rt.close()
    _reap(procs)


def test_per_gpu_nodes_register_separately(tmp_path):
    """``--per-gpu``: one process per listed device, each registering as its own node (on this CPU box the "devices" are just labels)."""
    from photon_b200.server.fleet import NodeFleetRuntime
    from photon_b200.server_app import run_server

    port = _free_port()
    procs = _spawn_nodes(1, port, extra=("--per-gpu", "--devices", "0,1"))
    cfg = _cfg(tmp_path, "run_uuid=pergpu", "photon.n_nodes=0", f"photon.fleet.address=127.0.0.1:{port}", "photon.fleet.n_remote_nodes=2",
               "photon.fleet.connect_timeout_s=120", "fl.n_rounds=1", "fl.n_clients_per_round=2", "llm_config.save_folder=null",
               "fl.strategy_name=fedavg", "photon.topology=nodes", "fl.eval_period=null")
    rt = NodeFleetRuntime(cfg)
    try:
        h = run_server(cfg, runtime=rt)
        assert sorted(rt.node_ids()) == [1000, 1001] and h.latest("server/n_failures") == 0
        assert sorted(s.info["devices"] for s in rt.link._slots.values()) == [[0], [1]]
    finally:
        rt.close()
    outs = _reap(procs)
    assert procs[0].returncode == 0, outs


def test_a_machine_can_join_a_running_federation(tmp_path):
    """The server waits for ONE remote node (``n_remote_nodes=1``); a second machine that connects later — here: whichever of two
    registers second — is admitted at the next health check and is part of the following rounds."""
    from photon_b200.server.fleet import NodeFleetRuntime
    from photon_b200.server_app import run_server

    port = _free_port()
    cfg = _cfg(tmp_path, "run_uuid=join", "photon.n_nodes=0", f"photon.fleet.address=127.0.0.1:{port}", "photon.fleet.n_remote_nodes=1",
               "photon.fleet.connect_timeout_s=120", "fl.n_rounds=6", "fl.n_clients_per_round=2", "llm_config.save_folder=null",
               "fl.strategy_name=fedavg", "photon.topology=nodes", "fl.eval_period=null")
    rt = NodeFleetRuntime(cfg)
    procs = _spawn_nodes(1, port)

    def late():
        t0 = time.time()
        while time.time() - t0 < 120 and (rt.link is None or not rt.link._slots):
            time.sleep(0.02)
        procs.extend(_spawn_nodes(1, port))          # the first node is in (the server is past its wait): a second machine shows up

    try:
        threading.Thread(target=late, daemon=True).start()
        orig = rt.node_ids

        def patient_node_ids():                       # keep the test independent of how fast a node process starts on this box
            if rt.link is not None and len(procs) == 2:
                t0 = time.time()
                while len(rt.link._slots) < 2 and time.time() - t0 < 90:
                    time.sleep(0.05)
            return orig()

        rt.node_ids = patient_node_ids
        h = run_server(cfg, runtime=rt)
        n_nodes = [v for _, v in h.metrics_distributed_fit["server/n_nodes"]]
        assert n_nodes[-1] == 2 and rt.n_remote_nodes == 1, n_nodes
        assert all(v == 0 for _, v in h.metrics_distributed_fit["server/n_failures"])
    finally:
        rt.close()
    outs = _reap(procs)
    assert all(p.returncode == 0 for p in procs), outs


@pytest.mark.parametrize("ranks", [1, 2])
def test_a_whole_box_joins_as_one_spmd_node(tmp_path, ranks):
    """``--spmd``: the box trains its share of the round with the SPMD runtime (one process per device, work queue, its own round
    transport for the box-level mean) and looks like one pre-aggregating node of capacity ``ranks`` to the server. Same global model
    as two ordinary in-process nodes."""
    from photon_b200.server.fleet import NodeFleetRuntime
    from photon_b200.server_app import run_server

    common = ["fl.n_rounds=2", "fl.n_clients_per_round=4", "llm_config.save_folder=null", "fl.strategy_name=fedavg", "photon.topology=nodes",
              "fl.eval_period=null"]
    want = _reference_model(tmp_path, "agg", common)
    common = common[:-1] + ["fl.eval_period=1"]          # the box also serves the federated evaluation
    port = _free_port()
    env = dict(os.environ, PYTHONPATH=str(ROOT), CUDA_VISIBLE_DEVICES="")
    node = [sys.executable, "-m", "photon_b200.node", "--server", f"127.0.0.1:{port}", "--spmd", "--max-idle-s", "60"]
    if ranks > 1:
        cmd = [sys.executable, "-m", "photon_b200.launch", "--nproc", str(ranks), "--master-port", str(_free_port()), "-m", "photon_b200.node", "--",
               *node[3:]]
    else:
        cmd = node
    proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    cfg = _cfg(tmp_path / f"box{ranks}", "run_uuid=agg", "photon.n_nodes=0", f"photon.fleet.address=127.0.0.1:{port}", "photon.fleet.n_remote_nodes=1",
               "photon.fleet.connect_timeout_s=180", *common)
    rt = NodeFleetRuntime(cfg)
    try:
        h = run_server(cfg, runtime=rt)
        assert [v for _, v in h.metrics_distributed_fit["server/n_failures"]] == [0, 0]
        assert len(h.losses_distributed) == 3 and all(2.0 < loss < 12.0 for _, loss in h.losses_distributed), h.losses_distributed
        node0 = rt.apps[0]
        assert node0.capacity == ranks and node0._slot.info["kind"] == "spmd-box"
        got = rt.round_backend.global_params().clone()
    finally:
        rt.close()
    out = _reap([proc])[0]
    assert proc.returncode == 0, out[-3000:]
    assert f"SPMD box, {ranks} rank(s)" in out
    assert torch.allclose(got, want, atol=1e-5), float((got - want).abs().max())


def test_node_rejoins_a_restarted_server(tmp_path):
    """The server process dies without saying goodbye and comes back on the same address (crash + resume): the node keeps its
    workers, notices that the link no longer knows it and registers again; a node pointed at ANOTHER run leaves instead."""
    from photon_b200.messages import Message
    from photon_b200.server.grpc_fleet import FleetLink

    cfg = _cfg(tmp_path, "run_uuid=phoenix", "llm_config.save_folder=null")
    port = _free_port()
    first = FleetLink(f"127.0.0.1:{port}", cfg=cfg, liveness_timeout_s=10)
    procs = _spawn_nodes(1, port)
    try:
        (node,) = first.wait_for_nodes(1, timeout_s=120)
        assert node.handle(Message("query", {"type": "free_resources"})).content == {"free_resources": {"status": "OK"}}
        first._server.stop(0)                                   # no shutdown message: a crash
        time.sleep(1.0)
        second = FleetLink(f"127.0.0.1:{port}", cfg=cfg, liveness_timeout_s=10)
        try:
            (again,) = second.wait_for_nodes(1, timeout_s=120)
            assert second._slots[again.node_id].info["rejoined_from"] == node.node_id and second._slots[again.node_id].info["pid"] == procs[0].pid
            assert again.handle(Message("query", {"type": "free_resources"})).content == {"free_resources": {"status": "OK"}}
            second._server.stop(0)
            time.sleep(1.0)
            other = FleetLink(f"127.0.0.1:{port}", cfg=_cfg(tmp_path, "run_uuid=someone-else", "llm_config.save_folder=null"))
            try:
                out = _reap(procs)[0]
                assert procs[0].returncode == 0 and "serves another run" in out and "re-registered as node" in out
            finally:
                other.close(grace_s=0.1)
        finally:
            second.close(grace_s=0.1)
    finally:
        first.close(grace_s=0.1)
        _reap(procs)


def test_an_idle_box_stays_together(tmp_path):
    """A two-rank box with a 2-second progress time-out sits idle for 7 seconds (the server has no work): rank 0 keeps ticking while it
    polls, the follower keeps waiting — and the next broadcast, a collective over both ranks, still goes through."""
    import numpy as np

    from photon_b200.clients.utils import get_initial_parameters
    from photon_b200.messages import Message
    from photon_b200.server.grpc_fleet import FleetLink

    cfg = _cfg(tmp_path, "run_uuid=idle", "llm_config.save_folder=null", "photon.progress_timeout_s=2", "photon.liveness_timeout_s=4")
    port = _free_port()
    link = FleetLink(f"127.0.0.1:{port}", cfg=cfg, liveness_timeout_s=30)
    env = dict(os.environ, PYTHONPATH=str(ROOT), CUDA_VISIBLE_DEVICES="")
    proc = subprocess.Popen([sys.executable, "-m", "photon_b200.launch", "--nproc", "2", "--master-port", str(_free_port()), "-m", "photon_b200.node", "--",
                             "--server", f"127.0.0.1:{port}", "--spmd", "--max-idle-s", "60"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        (node,) = link.wait_for_nodes(1, timeout_s=180)
        arrays, _ = get_initial_parameters(cfg)
        ok = {"broadcast": {"status": "OK"}}
        assert node.handle(Message("query", {"type": "broadcast_parameters", "parameters": [np.asarray(a) for a in arrays]})).content == ok
        time.sleep(7.0)
        assert node.alive()
        again = node.handle(Message("query", {"type": "broadcast_parameters", "parameters": [np.asarray(a) for a in arrays]}))
        assert again.content == ok and not again.error, again.error
    finally:
        link.close()
    out = _reap([proc])[0]
    assert proc.returncode == 0, out[-3000:]


def test_cross_host_example_script_runs_end_to_end(tmp_path):
    """``scripts/fed_cross_host_example.sh`` (resolver -> server with fleet link -> two SPMD boxes joining it) on a tiny model."""
    env = dict(os.environ, PYTHONPATH=str(ROOT), CUDA_VISIBLE_DEVICES="", PHOTON_STRICT_DATA="0", N_BOXES="2", N_ROUNDS="2", LOCAL_STEPS="1",
               N_CLIENTS="2", SAVE_PATH=str(tmp_path), RUN_UUID="xdemo", FLEET_PORT=str(_free_port()), N_GPUS="0",
               EXTERNAL_CONFIGS="llm_config.model.d_model=64 llm_config.model.n_heads=2 llm_config.model.n_layers=2 llm_config.max_seq_len=64 "
                                "llm_config.global_train_batch_size=4 llm_config.model.vocab_size=512")
    env.pop("PHOTON_SAVE_PATH", None)
    out = subprocess.run(["bash", str(ROOT / "scripts" / "fed_cross_host_example.sh")], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])
    assert "[server] done." in out.stdout and "'server/n_nodes': 2" in out.stdout and "'server/n_failures': 0" in out.stdout
    boxes = [(tmp_path / "xdemo" / f"box_{b}.log").read_text() for b in (0, 1)]
    assert all("SPMD box, 1 rank(s)" in b for b in boxes)


def test_fleet_link_over_tls(tmp_path):
    """``PHOTON_FLEET_TLS_KEY`` / ``_CERT`` on the server, ``--tls-ca`` on the node: the link is encrypted and the node checks the
    server's certificate (a self-signed one for ``localhost`` here); a node without the CA cannot talk to it."""
    import datetime as dt
    import ipaddress

    cryptography = pytest.importorskip("cryptography")
    del cryptography
    from cryptography import x509
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    from cryptography.x509.oid import NameOID

    from photon_b200.messages import Message
    from photon_b200.server.grpc_fleet import FleetLink

    key = ec.generate_private_key(ec.SECP256R1())
    name = x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, "localhost")])
    now = dt.datetime.now(dt.timezone.utc)
    cert = (x509.CertificateBuilder().subject_name(name).issuer_name(name).public_key(key.public_key()).serial_number(x509.random_serial_number())
            .not_valid_before(now - dt.timedelta(minutes=5)).not_valid_after(now + dt.timedelta(days=1))
            .add_extension(x509.SubjectAlternativeName([x509.DNSName("localhost"), x509.IPAddress(ipaddress.ip_address("127.0.0.1"))]), critical=False)
            .add_extension(x509.BasicConstraints(ca=True, path_length=None), critical=True).sign(key, hashes.SHA256()))
    kp, cp = tmp_path / "key.pem", tmp_path / "cert.pem"
    kp.write_bytes(key.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.PKCS8, serialization.NoEncryption()))
    cp.write_bytes(cert.public_bytes(serialization.Encoding.PEM))
    cfg = _cfg(tmp_path, "run_uuid=tls", "llm_config.save_folder=null")
    port = _free_port()
    link = FleetLink(f"127.0.0.1:{port}", cfg=cfg, token="t", tls=(str(kp), str(cp)))
    env = dict(os.environ, PYTHONPATH=str(ROOT), CUDA_VISIBLE_DEVICES="", PHOTON_FLEET_TOKEN="t")
    base = [sys.executable, "-m", "photon_b200.node", "--server", f"localhost:{port}", "--n-workers", "1", "--max-idle-s", "8"]
    good = subprocess.Popen(base + ["--tls-ca", str(cp)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        (node,) = link.wait_for_nodes(1, timeout_s=120)
        assert node.handle(Message("query", {"type": "free_resources"})).content == {"free_resources": {"status": "OK"}}
        plain = subprocess.Popen(base, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)      # no CA: plaintext to a TLS port
        out = _reap([plain])[0]
        assert plain.returncode != 0 and len(link._slots) == 1, out[-800:]
    finally:
        link.close()
    assert _reap([good]) and good.returncode == 0

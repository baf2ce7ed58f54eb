"""SPMD control plane on CPU (gloo ranks, host round transport): shared work queue, liveness, a SIGKILLed / hung rank mid-round.
The ranks are started by ``photon_b200.launch`` (torchrun would tear the whole job down with the first dead worker)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

from conftest import free_port as _free_port  # noqa: E402

ROOT = str(Path(__file__).resolve().parents[1])
TINY = ["llm_config.model.d_model=32", "llm_config.model.n_heads=2", "llm_config.model.n_layers=1", "llm_config.max_seq_len=16",
        "llm_config.global_train_batch_size=4", "llm_config.device_train_microbatch_size=2", "llm_config.device_eval_batch_size=4",
        "llm_config.local_steps=2ba", "llm_config.precision=fp32", "llm_config.model.attn_config.attn_impl=torch",
        "llm_config.log_to_console=false", "~llm_config.loggers.wandb", "~llm_config.loggers.tensorboard", "~llm_config.callbacks",
        "llm_config.save_folder=null", "fl.eval_period=null", "photon.resume_round=null", "photon.comm_stack.shm=true",
        "dataset.train.root_local=synthetic://c4", "dataset.val.root_local=synthetic://c4"]


def _launch(tmp_path, nproc, body, timeout=600):
    script = tmp_path / "run.py"
    script.write_text("import os, sys, torch, torch.distributed as dist\nsys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(body))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run([sys.executable, "-m", "photon_b200.launch", "--nproc", str(nproc), "--master-port", str(_free_port()), str(script)],
                          capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_work_queue_gives_the_fast_rank_more_clients(tmp_path):
    """6 clients, 2 ranks, rank 1 is a straggler (6 s per client): with the shared work queue it trains 1 client while
    the other takes the remaining 5 (a static client -> rank mapping would be 3 / 3); both ranks still end with the same model."""
    out = _launch(tmp_path, 2, f"""
        from photon_b200.config import compose
        from photon_b200.federation import FederationRuntime
        from photon_b200.server_app import run_server
        dist.init_process_group("gloo")
        cfg = compose({TINY!r} + ["run_uuid=wq", "fl.n_rounds=1", "fl.n_total_clients=6", "fl.n_clients_per_round=6",
                                "fl.fault_injection={{round: 1, rank: 1, kind: slow, seconds: 6}}"])
        rt = FederationRuntime(cfg, device=torch.device("cpu"), rank=dist.get_rank(), world_size=2)
        h = run_server(cfg, runtime=rt)
        if dist.get_rank() == 0:
            per_node = sorted(list(rt.assignment.values()).count(n) for n in (0, 1))
            assert per_node == [1, 5], rt.assignment
            assert h.metrics_distributed_fit["server/n_failures"][-1][1] == 0
            print("OK", rt.assignment)
    """)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.parametrize("kind", ["kill", "hang"])
def test_rank_lost_mid_round_the_federation_goes_on(tmp_path, kind):
    """3 ranks; rank 2 is SIGKILLed (or hangs forever) while training in round 2. The round completes over the survivors with the
    lost clients counted as failures (<= fl.accept_failures_cnt), the liveness check reports 2 nodes afterwards and round 3 runs on
    the 2 ranks that are left (the reference: photon/server_app.py:285,346 + node_manager_app.py:326-351,553-579)."""
    out = _launch(tmp_path, 3, f"""
        import math
        from photon_b200.config import compose
        from photon_b200.federation import FederationRuntime
        from photon_b200.server_app import run_server
        dist.init_process_group("gloo")
        cfg = compose({TINY!r} + ["run_uuid=ft", "fl.n_rounds=3", "fl.n_total_clients=6", "fl.n_clients_per_round=6", "fl.accept_failures_cnt=6",
                                "photon.liveness_timeout_s=3", "photon.progress_timeout_s=2",
                                "fl.fault_injection={{round: 2, rank: 2, kind: {kind}}}"])
        rt = FederationRuntime(cfg, device=torch.device("cpu"), rank=dist.get_rank(), world_size=3)
        h = run_server(cfg, runtime=rt)
        if dist.get_rank() == 0:
            fit = h.metrics_distributed_fit
            fails = dict(fit["server/n_failures"])
            assert fails[1] == 0 and fails[2] >= 1 and fails[3] == 0, fails
            assert "server/round_ignored" not in fit, "round 2 must aggregate the survivors, not be skipped"
            assert [r for r, _ in fit["server/l2_norm_pseudo_gradient"]] == [1, 2, 3]
            assert dict(h.metrics_centralized["server/n_nodes_after_round"])[2] == 2
            assert dict(fit["server/n_nodes"])[3] == 2
            assert sorted(set(rt.assignment.values())) == [0, 1]            # round 3 ran on the two survivors only
            assert math.isfinite(float(rt.round_backend.global_params().norm()))
            print("OK survivors", rt.alive_ranks)
        os._exit(0)      # no collective teardown: a peer is gone
    """)
    assert out.returncode == 0 and "OK survivors [0, 1]" in out.stdout, out.stdout[-2500:] + out.stderr[-3500:]


def test_control_plane_unit_gather_queue_and_dead_rank_in_one_process():
    """Three control planes (threads standing in for ranks) over one TCPStore: work-queue indices are unique and dense, ``gather`` is
    identical on every participant, a participant whose heartbeat stops is ruled out by rank 0 and everybody adopts that verdict."""
    import threading
    import time

    import torch

    from photon_b200.server.control import ControlPlane

    port = _free_port()
    planes = [None] * 3

    def make(r):
        planes[r] = ControlPlane(r, 3, host="127.0.0.1", port=port, heartbeat_s=0.1, liveness_timeout_s=1.0, namespace="unit", instance=1)

    ts = [threading.Thread(target=make, args=(r,)) for r in range(3)]
    [t.start() for t in ts], [t.join() for t in ts]
    assert all(p is not None for p in planes) and planes[0].alive() == [0, 1, 2]
    # work queue: every index exactly once
    q = [p.open_queue("round1") for p in planes]
    assert len(set(q)) == 1
    got = []
    lock = threading.Lock()

    def pull(p):
        while True:
            i = p.next_index(q[0])
            if i >= 10:
                return
            with lock:
                got.append(i)

    ts = [threading.Thread(target=pull, args=(p,)) for p in planes]
    [t.start() for t in ts], [t.join() for t in ts]
    assert sorted(got) == list(range(10))
    # gather / sum: same answer everywhere
    out = [None] * 3

    def ex(r):
        out[r] = (planes[r].gather("g", {"rank": r}), planes[r].sum_tensor("s", torch.tensor([float(r + 1)])))

    ts = [threading.Thread(target=ex, args=(r,)) for r in range(3)]
    [t.start() for t in ts], [t.join() for t in ts]
    assert all(sorted(o[0]) == [0, 1, 2] and float(o[1]) == 6.0 for o in out)
    # rank 2 "dies": its heartbeat stops; the next exchange completes over ranks 0 and 1 and both know why
    planes[2]._stop.set()
    time.sleep(1.3)

    def ex2(r):
        out[r] = planes[r].gather("after", r)

    ts = [threading.Thread(target=ex2, args=(r,)) for r in (0, 1)]
    [t.start() for t in ts], [t.join() for t in ts]
    assert sorted(out[0]) == [0, 1] and out[0] == out[1]
    assert planes[0].dead == {2} and planes[1].dead == {2} and planes[1].alive() == [0, 1]
    ts = [threading.Thread(target=p.close) for p in planes]
    [t.start() for t in ts], [t.join() for t in ts]


def test_two_ranks_with_federated_eval_and_server_checkpoints(tmp_path):
    """A 2-rank federation WITH evaluation rounds and server checkpoints: node 0 evaluates alone (its 1-rank Trainer must not
    all-reduce its metrics over the job's WORLD group — the other rank is not evaluating), FedAdam moments are checkpointed."""
    out = _launch(tmp_path, 2, f"""
        from photon_b200.config import compose
        from photon_b200.federation import FederationRuntime
        from photon_b200.server_app import run_server
        dist.init_process_group("gloo")
        cfg = compose({TINY!r} + ["run_uuid=ev", "fl.n_rounds=2", "fl.n_total_clients=4", "fl.n_clients_per_round=4", "fl.eval_period=1",
                                "llm_config.device_eval_batch_size=4", "llm_config.eval_subset_num_batches=1", "photon.checkpoint=true",
                                "photon.saving_path={tmp_path}", "fl.strategy_name=fedadam",
                                "fl.strategy_kwargs={{eta: 0.01, beta_1: 0.9, beta_2: 0.99, tau: 0.001}}"])
        rt = FederationRuntime(cfg, device=torch.device("cpu"), rank=dist.get_rank(), world_size=2)
        h = run_server(cfg, runtime=rt)
        if dist.get_rank() == 0:
            assert len(h.losses_distributed) == 3, h.losses_distributed        # rounds 0, 1, 2
            print("OK eval", [round(v, 3) for _, v in h.losses_distributed])
    """)
    assert out.returncode == 0 and "OK eval" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
    rounds = sorted(p.name for p in tmp_path.glob("*/ev/server/*") if p.is_dir())
    assert rounds == ["0", "1", "2"], rounds

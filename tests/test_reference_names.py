"""API-surface parity: every public top-level function / class name of the reference's ``photon`` package resolves under the
``photon`` import alias of this package (the list in ``fixtures/reference_public_names.json`` is names only, generated from the
reference tree with ``ast``), and the thin reference-shaped entry points added for that purpose actually work."""
import importlib
import json
import os
import queue
from pathlib import Path

import numpy as np
import pytest

import photon_b200  # noqa: F401 - installs the ``photon`` alias

NAMES = json.loads((Path(__file__).parent / "fixtures" / "reference_public_names.json").read_text())


@pytest.mark.parametrize("module", sorted(NAMES))
def test_every_public_reference_name_resolves(module):
    m = importlib.import_module(module)
    missing = [n for n in NAMES[module] if not hasattr(m, n)]
    assert not missing, f"{module}: {missing}"


def test_uri_checkpoint_helpers(tmp_path, monkeypatch):
    from photon.server.s3_utils import (NoCheckpointsFoundError, delete_clients_checkpoints, delete_object, delete_remote_object,
                                        delete_rounds, extract_s3_comm_config_from_configrecord, get_file_from_path,
                                        get_num_batches_from_checkpoint_name, list_objects, obtain_sorted_runs)
    from photon_b200.messages import ParamHandle

    monkeypatch.setenv("PHOTON_SAVE_PATH", str(tmp_path))
    for k in ("S3_ENDPOINT_URL",):
        monkeypatch.delenv(k, raising=False)
    run = tmp_path / "bkt" / "runA"
    keys = ("current_server_parameters", "current_momentum_vector")
    for r in (1, 2, 3):
        d = run / "server" / str(r)
        d.mkdir(parents=True)
        for k in keys:
            (d / f"{k}.npz").write_bytes(b"x")
        if r != 3:                      # round 3 is incomplete: no state.bin yet
            (d / "state.bin").write_bytes(b"s")
    for b in (10, 20, 30):
        p = run / "client_0" / f"ep0-ba{b}-rank0.pt"
        p.parent.mkdir(exist_ok=True)
        p.write_bytes(b"c")
    for uri in ("s3://bkt/runA", str(run)):          # the bucket as a URI (directory stand-in) and as a plain path
        assert obtain_sorted_runs(uri, keys) == [1, 2]
        ok, listed = list_objects(uri)
        assert ok and "server/1/state.bin" in listed and "client_0/ep0-ba20-rank0.pt" in listed
    with pytest.raises(NoCheckpointsFoundError):
        obtain_sorted_runs("s3://bkt/nothing-here", keys)
    assert get_num_batches_from_checkpoint_name("ep3-ba1280-rank0.pt") == 1280
    with pytest.raises(ValueError):
        get_num_batches_from_checkpoint_name("latest.pt")
    delete_rounds("s3://bkt/runA", keys)                                  # keeps the newest complete round
    assert obtain_sorted_runs("s3://bkt/runA", keys) == [2] and (run / "server" / "3").exists()
    delete_clients_checkpoints("s3://bkt/runA")                           # keeps each client's newest
    assert sorted(p.name for p in (run / "client_0").iterdir()) == ["ep0-ba30-rank0.pt"]
    delete_object("s3://bkt/runA/client_0/ep0-ba30-rank0.pt")
    delete_remote_object("s3://bkt/runA/server/3")
    assert list_objects("s3://bkt/runA/client_0") == (False, []) and not (run / "server" / "3" / "current_momentum_vector.npz").exists()
    f = run / "server" / "2" / "state.bin"
    assert get_file_from_path(str(f)) == f and get_file_from_path("s3://bkt/runA/server/2/state.bin", tmp_dir=tmp_path / "dl").read_bytes() == b"s"
    with pytest.raises(FileNotFoundError):
        get_file_from_path(str(run / "nope"))
    h = ParamHandle("file", "/x", {"endpoint_id": "node3", "folder_name": "comm_stack", "file_name": "parameters"})
    assert extract_s3_comm_config_from_configrecord(h) == ("node3", "parameters", "comm_stack")
    assert extract_s3_comm_config_from_configrecord({"endpoint_id": 7, "file_name": "p", "current_round": 4}) == ("7", "p", "4")


def test_reference_shaped_shm_helpers():
    from photon.shm.utils import (get_config_shm, get_dict_configsrecord_shm, get_ndarrays_size_and_bounds, get_num_samples_shm,
                                  is_shm_existing, set_config_shm, set_dict_configsrecord_shm, set_num_samples_shm)
    from photon_b200.shm.utils import unlink_quietly

    name = f"pb200_names_{os.getpid()}"
    try:
        total, bounds = get_ndarrays_size_and_bounds([np.zeros(3, np.float32), np.zeros((2, 2), np.float64)])
        assert total == 12 + 32 and bounds == [(0, 12), (12, 44)]
        view, shm = get_num_samples_shm(name + "_n", create=True)
        set_num_samples_shm(view, 42)
        again, shm2 = get_num_samples_shm(name + "_n")
        assert int(again[0]) == 42 and is_shm_existing(name + "_n")
        del view, again
        shm.close(), shm2.close()
        cfg, s1 = get_config_shm({"a": 1, "b": [1, 2]}, name + "_c", create=True)
        assert cfg == {"a": 1, "b": [1, 2]}
        set_config_shm({"a": 2}, s1)
        back, s2 = get_config_shm(None, name + "_c")
        assert back == {"a": 2}
        s1.close(), s2.close()
        rec, s3 = get_dict_configsrecord_shm(name + "_r", {"fit": {"lr": 0.1}}, create=True)
        set_dict_configsrecord_shm({"fit": {"lr": 0.2}}, s3)
        assert get_dict_configsrecord_shm(name + "_r")[0] == {"fit": {"lr": 0.2}} and rec == {"fit": {"lr": 0.1}}
        s3.close()
    finally:
        for suf in ("_n", "_c", "_r"):
            unlink_quietly(name + suf)
    assert not is_shm_existing(name + "_n")


def test_env_patcher_and_scheduler_generator_and_handlers():
    from photon.server.evaluate_utils import get_handle_success_and_failure_evaluate
    from photon.server.fit_utils import get_handle_success_and_failure_fit
    from photon.server.server_util import TooManyFailuresError, message_collaborative
    from photon.worker.utils import get_env_patcher
    from photon_b200.messages import Code, EvaluateRes, FitRes, Message, Status

    before = {k: os.environ.get(k) for k in ("RANK", "WORLD_SIZE", "MASTER_PORT", "RUN_NAME")}
    with get_env_patcher("run-x", "1", "29999", world_size=2):
        assert (os.environ["RANK"], os.environ["WORLD_SIZE"], os.environ["MASTER_PORT"], os.environ["RUN_NAME"]) == ("1", "2", "29999", "run-x")
    assert {k: os.environ.get(k) for k in before} == before

    # two "nodes", five sampled clients: each reply frees its node for the next client
    inbox: "queue.Queue[Message]" = queue.Queue()
    log: list[tuple[int, int]] = []

    def send(node, msg):
        cid = next(iter(msg.per_client))
        log.append((node, cid))
        inbox.put(Message("train", [FitRes(Status(Code.OK, ""), None, 8, {"cid": cid}, cid)], node_id=node, reply_to=msg.msg_id))

    def receive():
        out = []
        while not inbox.empty():
            out.append(inbox.get())
        return out

    replies = list(message_collaborative(send, receive, "train", [5, 6, 7, 8, 9], lambda rnd, cid: {"lr": 0.1 * cid}, [0, 1], 3, {}, 0, poll_s=0.0))
    assert sorted(r.content[0].cid for r in replies) == [5, 6, 7, 8, 9] and log[:2] == [(0, 5), (1, 6)] and {n for n, _ in log} == {0, 1}

    acc, failed = [], []
    h = get_handle_success_and_failure_fit(acc, failed, accept_failures_cnt=1)
    ok = FitRes(Status(Code.OK, ""), None, 4, {"loss": 1.0}, 0)
    assert h((True, ok)) == (True, ok) and acc == [({"loss": 1.0}, ok.status, 4)]
    assert h((False, None)) == (False, None) and failed == [None]
    with pytest.raises(TooManyFailuresError):
        h((False, None))
    eacc, efail = [], []
    he = get_handle_success_and_failure_evaluate(eacc, efail)
    er = EvaluateRes(Status(Code.OK, ""), 2.5, 10, {"ppl": 12.0})
    assert he((True, er))[0] and eacc[0][0] == 2.5 and not he((False, None))[0] and efail == [None]


def test_dataset_entry_points_and_broadcast_message():
    from photon.dataset.convert_dataset_hf import parse_args
    from photon.dataset.dataset_types import TokenizersCouple
    from photon.dataset.samples_generators import (generate_samples_from_dataloader, generate_samples_retokenized_streaming_text_dataset,
                                                    stream_and_untokenize)
    from photon.dataset.utils import build_dataloader, build_hf_dataset, check_tokenizer_config
    from photon.server.broadcast_utils import parameters_to_broadcast_recordset
    from photon_b200.dataset.utils import ByteTokenizer

    tok = ByteTokenizer()
    check_tokenizer_config(tok, eos_text="<|endoftext|>")
    loader = list(build_dataloader(build_hf_dataset("synthetic://20", "train", "CONCAT_TOKENS", None, 64, tokenizer=tok, eos_text="<|endoftext|>"), 4))
    assert loader[0]["tokens"].shape == (4, 64) and len(list(generate_samples_from_dataloader(loader, 5))) == 5
    text = list(stream_and_untokenize(loader, tok, truncate_num_batches=1))
    assert len(text) == 4 and all(isinstance(t, str) and t for t in text)
    re_tok = list(generate_samples_retokenized_streaming_text_dataset(loader, TokenizersCouple(tok, tok), 32, 7, no_wrap=False))
    assert len(re_tok) == 7 and re_tok[0]["tokens"].shape == (32,)
    assert len(list(build_hf_dataset("synthetic://3", "train", "NO_CONCAT"))) == 3
    assert parse_args(["--out_root", "/tmp/x", "--name", "it"]).dataset == "c4_it"
    arrays = [np.ones(3, np.float32)]
    msg = parameters_to_broadcast_recordset(arrays, keep_input=False)
    assert msg.kind == "query" and msg.content["type"] == "broadcast_parameters" and msg.content["parameters"].kind == "inline" and arrays == []


def test_client_app_module_callbacks(tmp_path):
    """``lifespan`` + the module-level ``query`` / ``train`` functions drive this process's node like the reference's Flower callbacks."""
    import photon.client_app as ca
    from photon_b200.clients.configs import get_photon_fit_config_fn
    from photon_b200.clients.utils import get_initial_parameters
    from photon_b200.messages import ClientState, Code, Message
    from photon_b200.server.server_util import fit_or_evaluate_ins
    from test_federation_cpu import _cfg

    cfg = _cfg(tmp_path, "run_uuid=cb", "llm_config.save_folder=null")
    with pytest.raises(RuntimeError, match="no node app"):
        ca.query(Message("query", {"type": "free_resources"}))
    arrays, _ = get_initial_parameters(cfg)
    with ca.lifespan(cfg, n_workers=1):
        ack = ca.set_parameters(Message("query", {"type": "broadcast_parameters", "parameters": arrays}))
        assert ack.content == {"broadcast": {"status": "OK"}}
        fn = get_photon_fit_config_fn(cfg)
        states = {c: ClientState() for c in range(4)}
        rep = ca.train(fit_or_evaluate_ins("train", 1, [2], states, 0, {2: fn(1, 2, states, 0).to_wire()}))
        assert [r.cid for r in rep.content] == [2] and rep.content[0].status.code == Code.OK
        assert ca.free_resources().content == {"free_resources": {"status": "OK"}}
    with pytest.raises(RuntimeError):
        ca.evaluate(Message("evaluate", None))

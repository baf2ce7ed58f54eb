"""Shard format, streaming dataset, shm codecs, checkpoint store, client load-path table — CPU."""
import numpy as np
import pytest
import torch

from photon_b200.checkpoint import CheckpointStore
from photon_b200.clients.llm_config_functions import (client_set_data_config, get_train_config, set_client_load_path,
                                                      set_dataset_default_params)
from photon_b200.config import compose
from photon_b200.data.shards import ShardReader, ShardWriter
from photon_b200.data.streaming import Stream, StreamingTokenDataset, TokenLoader, build_text_loader
from photon_b200.data.synthetic import SyntheticC4
from photon_b200.shm import ModelParametersMetadata, close_all_shms, get_parameters_shm, set_parameters_shm, shm_exists
from photon_b200.utils.core import dump_model_parameters_to_file, load_model_parameters_from_file, parameters_checker
from photon_b200.utils.flat import FlatLayout


@pytest.mark.parametrize("comp", [None, "zlib", "zstd"])
def test_shard_roundtrip(tmp_path, comp):
    rows = [np.arange(i, i + 16, dtype=np.int32) for i in range(37)]
    with ShardWriter(tmp_path / "s", seq_len=16, shard_samples=10, compression=comp) as w:
        w.write_many(rows)
    r = ShardReader(tmp_path / "s", validate_hash=True)
    assert len(r) == 37 and len(r.index["shards"]) == 4
    assert all(np.array_equal(r[i], rows[i]) for i in (0, 9, 10, 36))
    with pytest.raises(IndexError):
        r[37]


def test_synthetic_deterministic_and_c4_shaped():
    a, b = SyntheticC4(seed=1, stream_id=2)[5], SyntheticC4(seed=1, stream_id=2)[5]
    assert np.array_equal(a, b) and a.dtype == np.int32 and a.shape == (2048,) and a.max() < 50277
    assert (a == 0).sum() >= 1 and not np.array_equal(a, SyntheticC4(seed=1, stream_id=3)[5])
    assert abs(SyntheticC4().unigram_probabilities().sum() - 1.0) < 1e-9


def test_streaming_partition_resume_and_mixing(tmp_path):
    for c in range(2):
        with ShardWriter(tmp_path / f"c{c}" / "train", seq_len=8, shard_samples=16) as w:
            w.write_many([np.full(8, 100 * c + i, dtype=np.int32) for i in range(20)])
    streams = [Stream(local=str(tmp_path / "c0"), split="train"), Stream(local=str(tmp_path / "c1"), split="train", choose=10)]
    seen = []
    for rank in range(2):
        ds = StreamingTokenDataset(streams, seq_len=8, rank=rank, world_size=2, shuffle=True, allow_synthetic=False)
        seen.append([int(x[0]) for x in ds])
    assert len(seen[0]) == len(seen[1]) == 15 and not set(seen[0]) & set(seen[1])
    ds = StreamingTokenDataset(streams, seq_len=8, shuffle=True)
    it = iter(ds)
    first = [int(next(it)[0]) for _ in range(7)]
    sd = ds.state_dict()
    rest = [int(x[0]) for x in it]
    ds2 = StreamingTokenDataset(streams, seq_len=8, shuffle=True)
    ds2.load_state_dict(sd)
    assert [int(x[0]) for x in ds2] == rest and len(first) + len(rest) == 30
    ld = TokenLoader(StreamingTokenDataset(streams, seq_len=8), batch_size=4, num_workers=1)
    batches = list(ld)
    assert len(batches) == len(ld) == 7 and batches[0]["input_ids"].shape == (4, 8) and batches[0]["input_ids"].dtype == torch.int64


def test_client_stream_selection_from_config():
    cfg = compose(["dataset.train.root_local=/data/x", "dataset/streams@dataset.val.streams=8_clients"])
    t = get_train_config(cfg.llm_config, run_uuid="r")
    evals = client_set_data_config(t, cid=10, split_eval=True)
    st = t.train_loader.dataset.streams
    assert list(st) == ["stream_2"] and st["stream_2"]["local"] == "/data/x/c8/en/client_2" and st["stream_2"]["split"] == "train"   # 10 % 8
    assert sorted(evals) == [f"client_{i}" for i in range(8)]
    t2 = get_train_config(cfg.llm_config)
    client_set_data_config(t2, cid=None)
    assert len(t2.train_loader.dataset.streams) == 8                                           # cid=None concatenates all
    set_dataset_default_params(t2)
    d = t2.train_loader.dataset
    assert d.predownload == 8 * 256 and d.num_canonical_nodes == 64 and d.shuffle_block_size == 1 << 18
    ld = build_text_loader({"dataset": {"streams": {"s": {"local": "synthetic://7"}}, "max_seq_len": 32}, "drop_last": True}, 2)
    assert next(iter(ld))["input_ids"].shape == (2, 32)


def test_shm_metadata_and_views():
    arrays = [np.random.randn(5, 3).astype(np.float32), np.arange(7, dtype=np.float32)]
    shm, meta = set_parameters_shm("pb200_ut_par", arrays)
    m2 = ModelParametersMetadata.from_literal(meta.to_literal())
    assert m2.same_layout(meta) and m2.total_num_bytes == meta.total_num_bytes
    h, views = get_parameters_shm("pb200_ut_par", m2)
    parameters_checker(arrays, views, expect_equal=True)
    views[1][0] = 42.0                                                    # zero-copy: visible through the writer's handle
    assert np.ndarray((7,), dtype=np.float32, buffer=shm.buf[meta.array_bounds[1][0]:meta.array_bounds[1][1]])[0] == 42.0
    with pytest.raises(AssertionError):
        parameters_checker(arrays, views, expect_equal=True)
    del views
    h.close(), shm.close()
    close_all_shms("pb200_ut_par")
    assert not shm_exists("pb200_ut_par")


def test_npz_sorted_order_and_checkpoint_store(tmp_path):
    lay = FlatLayout.build([("transformer.blocks.10.w", (2, 2)), ("transformer.blocks.2.w", (3,)), ("transformer.a", (1,))], align=4, total_multiple=4)
    assert lay.names == ("transformer.a", "transformer.blocks.10.w", "transformer.blocks.2.w")
    flat = torch.arange(lay.total, dtype=torch.float32)
    p = dump_model_parameters_to_file(tmp_path / "m.npz", lay.to_ndarrays(flat))
    with np.load(p) as z:
        assert z.files == ["arr_0", "arr_1", "arr_2"] and z["arr_1"].shape == (2, 2)
    assert [a.shape for a in load_model_parameters_from_file(p)] == [(1,), (2, 2), (3,)]
    store = CheckpointStore(tmp_path, "checkpoints")
    keys = ["current_server_parameters", "current_momentum_vector"]
    for r in (0, 1, 2):
        store.upload_server_checkpoint("runA", r, layout=lay, tensors={k: flat + r for k in keys}, state={"client_state": "{}", "server_steps_cumulative": r * 8})
    (store.round_dir("runA", 3)).mkdir()                                    # incomplete round: ignored
    assert store.obtain_sorted_rounds("runA", keys) == [0, 1, 2]
    assert store.interpret_resume_round("runA", -1, keys) == 2 and store.interpret_resume_round("runA", None, keys) is None
    with pytest.raises(FileNotFoundError):
        store.interpret_resume_round("runA", 3, keys)
    tensors, state = store.download_server_checkpoint("runA", 1, layout=lay, state_keys=keys)
    assert all(torch.equal(lay.view(tensors[keys[1]], i), lay.view(flat + 1, i)) for i in range(3)) and state["server_steps_cumulative"] == 8
    store.copy_old_checkpoints_to_new_run("runA", "runB", 2, state_keys=keys)
    assert store.obtain_sorted_rounds("runB", keys) == [2]
    store.cleanup_checkpoints("runA", per_round=True)
    assert store.obtain_sorted_rounds("runA", keys) == [2]
    store.cleanup_checkpoints("runA")
    assert store.list_objects("runA") == []


def test_set_client_load_path_decision_table(tmp_path):
    def cfg():
        return get_train_config({"save_folder": str(tmp_path / "ck"), "run_name": "r"})

    t = cfg()
    assert set_client_load_path(t, 3, 128) == (False, False) and t.save_folder.endswith("client_3")
    d = tmp_path / "ck" / "client_3"
    d.mkdir(parents=True)
    (d / "ep0-ba64-rank0.pt").write_bytes(b"x")
    t = cfg()
    assert set_client_load_path(t, 3, 128) == (False, True) and t.load_path.endswith("ep0-ba64-rank{rank}.pt")   # resume mid-round
    (d / "ep0-ba128-rank0.pt").write_bytes(b"x")
    t = cfg()
    assert set_client_load_path(t, 3, 128) == (True, True) and t.save_folder is None                             # round already done: skip
    t = cfg()
    assert set_client_load_path(t, 3, 32) == (False, False)                                                     # only newer checkpoints


def test_object_store_facade_and_run_listing(tmp_path):
    """RemoteUploaderDownloader call surface (ref: photon/utils.py:955-1014) and CheckpointStore.obtain_sorted_runs."""
    from photon_b200.checkpoint.store import CheckpointStore
    from photon_b200.utils.core import create_remote_up_down

    r = create_remote_up_down("bkt", "pre", "run1", 2, root=tmp_path)
    src = tmp_path / "a.bin"
    src.write_bytes(b"hello")
    r.upload_file("x/y/a.bin", src)
    assert r.list_objects() == ["x/y/a.bin"] and r.list_objects("x") == ["x/y/a.bin"]
    with pytest.raises(FileExistsError):
        r.upload_file("x/y/a.bin", src, overwrite=False)
    r.download_file("x/y/a.bin", tmp_path / "out" / "b.bin")
    assert (tmp_path / "out" / "b.bin").read_bytes() == b"hello"
    r.delete_object("x/y/a.bin")
    assert r.list_objects() == []

    store = CheckpointStore(tmp_path / "ck")
    for run in ("r_old", "r_new"):
        (store.server_dir(run) / "1").mkdir(parents=True)
    import os
    import time
    os.utime(store.bucket / "r_old", (time.time() - 100, time.time() - 100))
    assert store.obtain_sorted_runs() == ["r_old", "r_new"]


def test_streaming_shm_cleanup_removes_only_our_segments():
    from multiprocessing import shared_memory

    from photon_b200.clients.utils import streaming_shms_clean_up

    ours = shared_memory.SharedMemory(name="pb200_test_leak", create=True, size=64)
    other = shared_memory.SharedMemory(name="zz_not_ours_seg", create=True, size=64)
    try:
        ours.close()
        assert streaming_shms_clean_up() >= 1
        with pytest.raises(FileNotFoundError):
            shared_memory.SharedMemory(name="pb200_test_leak")
        shared_memory.SharedMemory(name="zz_not_ours_seg").close()   # untouched
    finally:
        other.close()
        other.unlink()


def test_loader_state_tracks_the_consumer_not_the_prefetch_thread():
    """With a prefetch thread the dataset cursor runs ahead of what was trained on; the checkpointed position must be the
    consumer's, so a resumed loader continues with exactly the next unseen batch."""
    from photon_b200.data.streaming import Stream, StreamingTokenDataset, TokenLoader

    def make():
        ds = StreamingTokenDataset([Stream(local="synthetic://7", name="s")], seq_len=8, shuffle=True, shuffle_seed=3, synthetic_samples=64)
        return TokenLoader(ds, batch_size=4, num_workers=1, prefetch=4, pin_memory=False)

    ref = [b["input_ids"].clone() for b in make()]
    assert len(ref) == 16
    a = make()
    it = iter(a)
    seen = [next(it)["input_ids"].clone() for _ in range(5)]
    import time

    t0 = time.time()
    while a.dataset.sample_in_epoch <= 20 and time.time() - t0 < 10:   # let the prefetcher run ahead
        time.sleep(0.02)
    sd = a.state_dict()
    assert sd["sample_in_epoch"] == 20 and a.dataset.sample_in_epoch > 20
    b = make()
    b.load_state_dict(sd)
    rest = [x["input_ids"] for x in b]
    assert len(rest) == 11 and all(torch.equal(x, y) for x, y in zip(seen + rest, ref))
    it.close()
    assert a.dataset.sample_in_epoch == 20 and a.state_dict()["sample_in_epoch"] == 20   # abandoned iterator: prefetched samples are not lost


def test_state_bin_written_by_the_reference_is_readable(tmp_path, monkeypatch):
    """The reference pickles a Flower ``History`` subclass into ``state.bin``; neither ``flwr`` nor ``photon`` is importable
    here, so the loader must read it without those classes and hand back our own history object."""
    import importlib
    import pickle
    import sys

    from photon_b200.checkpoint.store import load_server_state
    from photon_b200.messages import decode_client_states
    from photon_b200.wandb_history import WandbHistory

    pkg = tmp_path / "fakeflwr_pkg"
    pkg.mkdir()
    (pkg / "fakeflwr_history.py").write_text(
        "class History:\n"
        "    def __init__(self):\n"
        "        self.losses_distributed, self.losses_centralized = [(0, 11.0), (1, 10.5)], []\n"
        "        self.metrics_distributed_fit, self.metrics_distributed, self.metrics_centralized = {'server/n_failures': [(1, 0)]}, {}, {'t': [(1, 2.0)]}\n"
        "class WandbHistory(History):\n"
        "    def __init__(self):\n"
        "        super().__init__()\n"
        "        self.use_wandb = False\n")
    monkeypatch.syspath_prepend(str(pkg))
    mod = importlib.import_module("fakeflwr_history")
    blob = pickle.dumps({"server_round": 1, "history": mod.WandbHistory(), "time_offset": 3.5, "server_steps_cumulative": 4,
                         "client_state": "{0: {'local_steps_cumulative': 4, 'local_timestamp': {}, 'steps_done': 0}}"})
    del sys.modules["fakeflwr_history"]
    monkeypatch.setattr(sys, "path", [p for p in sys.path if p != str(pkg)])
    importlib.invalidate_caches()
    (tmp_path / "state.bin").write_bytes(blob)
    st = load_server_state(tmp_path / "state.bin")
    assert isinstance(st["history"], WandbHistory) and st["history"].losses_distributed == [(0, 11.0), (1, 10.5)]
    assert st["history"].latest("server/n_failures") == 0 and st["server_steps_cumulative"] == 4 and st["time_offset"] == 3.5
    assert decode_client_states(st["client_state"])[0].local_steps_cumulative == 4


def test_mds_bytes_column_of_older_converters(tmp_path):
    """``columns={"tokens": "bytes"}`` with int64 ids (older llm-foundry ``ConcatTokensDataset``)."""
    import json as _json

    from photon_b200.data.shards import MDSReader

    rows = [np.arange(i, i + 16, dtype=np.int64) for i in range(3)]
    cols = {"column_encodings": ["bytes"], "column_names": ["tokens"], "column_sizes": [None]}
    config = _json.dumps({**cols, "format": "mds"}).encode()
    samples = [np.uint32(r.nbytes).tobytes() + r.tobytes() for r in rows]
    offsets = np.array([0] + [len(x) for x in samples]).cumsum().astype(np.uint32)
    offsets += 4 + offsets.nbytes + len(config)
    d = tmp_path / "old"
    d.mkdir()
    (d / "shard.00000.mds").write_bytes(np.uint32(3).tobytes() + offsets.tobytes() + config + b"".join(samples))
    (d / "index.json").write_text(_json.dumps({"version": 2, "shards": [{**cols, "compression": None, "format": "mds", "samples": 3, "version": 2,
                                                                        "raw_data": {"basename": "shard.00000.mds", "bytes": 0, "hashes": {}}, "zip_data": None}]}))
    r = MDSReader(d)
    assert r.seq_len == 16 and np.array_equal(r[2], rows[2].astype(np.int32)) and r[0].dtype == np.int32


def _write_mds(directory, rows, shard_samples, compression=None):
    """mosaicml-streaming's MDS layout for ``columns={"tokens": "ndarray:int32"}`` (see MDSReader's docstring)."""
    import json as _json

    import pyarrow as pa

    directory.mkdir(parents=True)
    cols = {"column_encodings": ["ndarray:int32"], "column_names": ["tokens"], "column_sizes": [None]}
    config = _json.dumps({**cols, "compression": compression, "format": "mds", "hashes": [], "size_limit": 1 << 26}, sort_keys=True).encode()
    shards = []
    for k in range(0, len(rows), shard_samples):
        samples = []
        for r in rows[k:k + shard_samples]:
            payload = bytes([(1 << 4) | 1]) + int(r.size).to_bytes(2, "little") + r.astype(np.int32).tobytes()   # rank-1, uint16 dim, data
            samples.append(np.uint32(len(payload)).tobytes() + payload)
        n = np.uint32(len(samples))
        offsets = np.array([0] + [len(x) for x in samples]).cumsum().astype(np.uint32)
        offsets += 4 + offsets.nbytes + len(config)
        raw = n.tobytes() + offsets.tobytes() + config + b"".join(samples)
        base = f"shard.{len(shards):05d}.mds"
        meta = {**cols, "compression": compression, "format": "mds", "hashes": [], "samples": len(samples), "size_limit": 1 << 26, "version": 2,
                "raw_data": {"basename": base, "bytes": len(raw), "hashes": {}}, "zip_data": None}
        if compression:
            z = pa.Codec("zstd").compress(raw, asbytes=True)
            (directory / (base + ".zstd")).write_bytes(z)
            meta["zip_data"] = {"basename": base + ".zstd", "bytes": len(z), "hashes": {}}
        else:
            (directory / base).write_bytes(raw)
        shards.append(meta)
    (directory / "index.json").write_text(_json.dumps({"shards": shards, "version": 2}))


@pytest.mark.parametrize("compression", [None, "zstd"])
def test_mds_directories_written_for_the_reference_are_readable(tmp_path, compression):
    from photon_b200.data.shards import MDSReader, open_shard_dir
    from photon_b200.data.streaming import Stream, StreamingTokenDataset, TokenLoader

    rng = np.random.default_rng(0)
    rows = [rng.integers(0, 50000, 32).astype(np.int32) for _ in range(25)]
    _write_mds(tmp_path / "c8" / "en" / "client_0" / "train", rows, shard_samples=10, compression=compression)
    r = open_shard_dir(tmp_path / "c8" / "en" / "client_0" / "train")
    assert isinstance(r, MDSReader) and len(r) == 25 and r.seq_len == 32
    assert all(np.array_equal(r[i], rows[i]) for i in (0, 9, 10, 24))
    ds = StreamingTokenDataset([Stream(local=str(tmp_path / "c8" / "en" / "client_0"), split="train", name="s")], seq_len=32, shuffle=False,
                               allow_synthetic=False)
    batches = [b["input_ids"] for b in TokenLoader(ds, batch_size=5, pin_memory=False)]
    assert len(batches) == 5 and torch.equal(batches[0][0], torch.from_numpy(rows[0].astype(np.int64)))


import pathlib

FIXTURES = pathlib.Path(__file__).resolve().parent / "fixtures"


def test_committed_mds_fixture_opens_and_streams():
    """A mosaicml-streaming MDS directory (index.json v2 + zstd shards, written by tests/fixtures/make_fixtures.py byte by byte from
    the format, independently of the reader) is readable as-is: the reference's converted datasets need no re-tokenising
    (ref: photon/dataset/convert_dataset_hf.py:323-327)."""
    import numpy as np

    from photon_b200.data.shards import MDSReader, open_shard_dir
    from photon_b200.data.streaming import Stream, StreamingTokenDataset

    r = open_shard_dir(FIXTURES / "mds_tiny")
    assert isinstance(r, MDSReader) and len(r) == 12 and r.seq_len == 16
    want = np.load(FIXTURES / "mds_tiny" / "expected_tokens.npy")
    for i in range(12):
        assert r[i].dtype == np.int32 and (r[i] == want[i]).all()
    ds = StreamingTokenDataset([Stream(local=str(FIXTURES / "mds_tiny"), split=None, name="mds")], seq_len=16, allow_synthetic=False)
    got = np.stack([np.asarray(x) for _, x in zip(range(12), iter(ds))])
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))


def test_committed_reference_state_bin_loads():
    """A ``state.bin`` carrying the reference's pickled ``photon.wandb_history.WandbHistory`` (Flower ``History`` subclass) and its
    five fields loads without flwr: history stores, literal client-state string, counters (ref: photon/server/s3_utils.py:374-389)."""
    import ast

    from photon_b200.checkpoint.store import load_server_state
    from photon_b200.wandb_history import History

    st = load_server_state(FIXTURES / "ref_state" / "state.bin")
    assert st["server_round"] == 2 and st["server_steps_cumulative"] == 256 and abs(st["time_offset"] - 1234.5) < 1e-9
    h = st["history"]
    assert isinstance(h, History)
    assert h.losses_distributed == [(0, 10.83), (1, 9.91), (2, 9.17)]
    assert h.metrics_distributed_fit["server/n_aggregated_clients"] == [(1, 8), (2, 8)]
    cs = ast.literal_eval(st["client_state"])
    assert set(cs) == set(range(8)) and cs[3]["steps_done"] == 128


def test_background_checkpoint_writer_snapshots_now_and_writes_in_order(tmp_path):
    """``background=True``: the planes and the state are captured at the call (later mutations do not leak into the file), writes
    land in submission order, ``wait()`` makes them visible and re-raises a writer error."""
    import numpy as np
    import torch

    from photon_b200.checkpoint.store import CheckpointStore, load_server_state
    from photon_b200.utils.flat import FlatLayout

    lay = FlatLayout.build([("a", (300,)), ("b", (5, 7))])
    store = CheckpointStore(tmp_path, "bucket")
    flat = torch.arange(lay.total, dtype=torch.float32)
    hist = {"rounds": [1]}
    for r in (1, 2, 3):
        store.upload_server_checkpoint("run", r, layout=lay, tensors={"current_server_parameters": flat}, background=True,
                                       state={"client_state": "{}", "server_steps_cumulative": r, "seen_rounds": hist})
        flat += 1000.0            # the next round's update happens while the writer is still busy
        hist["rounds"].append(r + 1)
    store.wait()
    assert store.obtain_sorted_rounds("run", ["current_server_parameters"]) == [1, 2, 3]
    for r in (1, 2, 3):
        with np.load(store.round_dir("run", r) / "current_server_parameters.npz") as z:
            assert float(z["arr_0"][1]) == 1.0 + 1000.0 * (r - 1)         # the value AT round r, not the final one
        st = load_server_state(store.round_dir("run", r) / "state.bin")
        assert st["server_steps_cumulative"] == r and st["seen_rounds"]["rounds"] == list(range(1, r + 1))
    store.submit(lambda: (_ for _ in ()).throw(OSError("disk full")))
    import pytest

    with pytest.raises(RuntimeError, match="background checkpoint write failed"):
        store.wait()

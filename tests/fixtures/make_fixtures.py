"""Writes the two interop fixtures committed next to this file (run once; the outputs are tracked):

* ``mds_tiny/``  a mosaicml-streaming MDS directory exactly as ``MDSWriter(columns={"tokens": "ndarray:int32"},
  compression="zstd")`` lays it out (the reference's dataset format, ref: photon/dataset/convert_dataset_hf.py:323-327):
  ``index.json`` (version 2) + ``shard.0000N.mds.zstd``; a raw shard is
  ``uint32 n | uint32 offsets[n+1] | json config | samples``, a sample is ``uint32 sizes of the variable-size columns | payloads``,
  an ``ndarray:int32`` payload is ``shape header | C-order data``. The shape header is written here in its ONE-byte form
  (``uint8 ndim | dims``); ``photon_b200.data.shards.MDSWriter`` writes the two-byte form (``uint8 ndim | uint8 dims-type | dims``)
  that mosaicml-streaming's NDArray encoding is believed to use — no install was available to settle it, so the reader accepts
  both and always checks the dims against the payload length. Written byte by byte, independently of the reader under test.
* ``ref_state/state.bin``  a server-state pickle with the reference's five fields and a ``photon.wandb_history.WandbHistory``
  (a Flower ``History`` subclass) object inside (ref: photon/server/s3_utils.py:374-389), produced with stand-in classes that
  carry the reference's module / class names so the byte stream is what the reference's ``pickle.dump`` emits.
"""
from __future__ import annotations

import json
import pickle
import sys
import types
from pathlib import Path

import numpy as np
import pyarrow as pa

HERE = Path(__file__).resolve().parent


def ndarray_int32_payload(a: np.ndarray) -> bytes:
    # encoding "ndarray:int32" (dtype fixed by the encoding, shape dynamic): 1 header byte = ndim, then the dims in the smallest
    # unsigned type that holds them (uint8 here: 16 tokens), then the data
    assert a.dtype == np.int32 and a.ndim == 1 and a.shape[0] < 256
    return np.uint8(a.ndim).tobytes() + np.array(a.shape, np.uint8).tobytes() + a.tobytes()


def mds_shard(samples: list[np.ndarray], config: dict) -> bytes:
    enc = []
    for a in samples:
        body = ndarray_int32_payload(a)
        enc.append(np.array([len(body)], np.uint32).tobytes() + body)       # head: sizes of the variable-size columns, then bodies
    cfg = json.dumps(config, sort_keys=True).encode()
    n = np.uint32(len(enc))
    offsets = np.array([0] + [len(e) for e in enc]).cumsum().astype(np.uint32)
    offsets += np.uint32(4 + offsets.nbytes + len(cfg))
    return n.tobytes() + offsets.tobytes() + cfg + b"".join(enc)


def write_mds(out: Path) -> None:
    out.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(7)
    codec = pa.Codec("zstd", compression_level=3)
    config = {"version": 2, "format": "mds", "compression": "zstd", "hashes": [], "size_limit": 1 << 26,
              "column_names": ["tokens"], "column_encodings": ["ndarray:int32"], "column_sizes": [None]}
    shards, all_rows = [], []
    for s, count in enumerate((7, 5)):
        rows = [rng.integers(0, 50277, 16).astype(np.int32) for _ in range(count)]
        all_rows += rows
        raw = mds_shard(rows, config)
        base = f"shard.{s:05d}.mds"
        z = codec.compress(raw, asbytes=True)
        (out / (base + ".zstd")).write_bytes(z)
        shards.append({**config, "samples": count, "raw_data": {"basename": base, "bytes": len(raw), "hashes": {}},
                       "zip_data": {"basename": base + ".zstd", "bytes": len(z), "hashes": {}}})
    (out / "index.json").write_text(json.dumps({"version": 2, "shards": shards}, sort_keys=True))
    np.save(out / "expected_tokens.npy", np.stack(all_rows))


def write_state(out: Path) -> None:
    out.mkdir(parents=True, exist_ok=True)
    saved = {k: sys.modules.get(k) for k in ("flwr", "flwr.server", "flwr.server.history", "photon", "photon.wandb_history")}
    try:
        hist_mod = types.ModuleType("flwr.server.history")

        class History:  # flwr.server.history.History's attribute set
            def __init__(self) -> None:
                self.losses_distributed, self.losses_centralized = [], []
                self.metrics_distributed_fit, self.metrics_distributed, self.metrics_centralized = {}, {}, {}

        History.__module__, History.__qualname__ = "flwr.server.history", "History"
        hist_mod.History = History
        wh_mod = types.ModuleType("photon.wandb_history")

        class WandbHistory(History):
            def __init__(self, use_wandb: bool = True) -> None:
                super().__init__()
                self.use_wandb = use_wandb

        WandbHistory.__module__, WandbHistory.__qualname__ = "photon.wandb_history", "WandbHistory"
        wh_mod.WandbHistory = WandbHistory
        for name, mod in (("flwr", types.ModuleType("flwr")), ("flwr.server", types.ModuleType("flwr.server")), ("flwr.server.history", hist_mod),
                          ("photon", types.ModuleType("photon")), ("photon.wandb_history", wh_mod)):
            sys.modules[name] = mod
        h = WandbHistory(False)
        h.losses_distributed = [(0, 10.83), (1, 9.91), (2, 9.17)]
        h.metrics_distributed_fit = {"server/l2_norm_pseudo_gradient": [(1, 3.25), (2, 2.5)], "server/n_aggregated_clients": [(1, 8), (2, 8)]}
        h.metrics_distributed = {"ValLanguagePerplexity": [(0, 50500.0), (1, 20100.0), (2, 9600.0)]}
        client_state = {c: {"local_steps_cumulative": 256, "local_timestamp": {"batch": 256, "epoch": 0}, "steps_done": 128} for c in range(8)}
        state = {"server_round": 2, "history": h, "time_offset": 1234.5, "client_state": str(client_state), "server_steps_cumulative": 256}
        (out / "state.bin").write_bytes(pickle.dumps(state, protocol=4))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


if __name__ == "__main__":
    write_mds(HERE / "mds_tiny")
    write_state(HERE / "ref_state")
    print("fixtures written under", HERE)

"""Parameter side channel + the server-checkpoint entry points under the reference's module name
(ref: photon/server/s3_utils.py — sender :730-933, receiver :936-1115, checkpoint API :215-727,1261-1641).

The reference never ships tensors inside a control message: the sender parks the payload somewhere both
sides can reach (S3 object, POSIX shm segment, Ray object store) and the message carries only a locator.
Here the message field is a ``ParamHandle`` and the places are

============  ==========================================================================================
``nvl``       the symmetric NVLink arena — nothing to move; the fused round kernel reads/writes the planes
``inline``    same process (SPMD runtime, tests): the flat tensor / ndarray list itself
``shm``       one flat POSIX segment named ``endpoint_id`` + ``ModelParametersMetadata`` (node manager path)
``file``      ``{root}/{folder_name}/{endpoint_id}/{file_name}.npz`` — a directory standing in for the bucket
``s3``        key of an npz object in the configured S3 bucket (``utils/objstore.py``); the way parameters reach nodes on
              OTHER hosts (``server/grpc_fleet.py``)
``link``      key of an npz object in the fleet link's own object service (no S3 configured; the reference's Ray object store role)
============  ==========================================================================================

``replace_remote_with_parameters_in_recordset`` (sender) and ``replace_parameters_in_recordset_with_remote``
(receiver) are the two halves; they keep the reference's names although there is no RecordSet any more.
The checkpoint functions live in ``photon_b200.checkpoint.store.CheckpointStore`` and are re-exported.
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Any, Sequence

import numpy as np
import torch

from photon_b200.checkpoint.store import CheckpointStore, load_pretrained_model_from_path  # noqa: F401 - re-export
from photon_b200.messages import ParamHandle
from photon_b200.shm.utils import (ModelParametersMetadata, get_parameters_shm, set_parameters_shm, unlink_quietly)
from photon_b200.utils.core import dump_model_parameters_to_file, load_model_parameters_from_file
from photon_b200.utils.flat import FlatLayout

_LIVE_SEGMENTS: dict[str, Any] = {}  # sender keeps its segments mapped until ``release_remote_parameters``
_LINK_STORE: Any = None              # the fleet link's object service in this process (server: its spool, node: a LinkObjectStore)


def set_link_store(store: Any) -> None:
    global _LINK_STORE
    _LINK_STORE = store


def get_link_store() -> Any:
    return _LINK_STORE


def _comm_kind(comm_stack: Any) -> str:
    """First enabled transport of a ``photon.comm_stack`` node, in the reference's priority order."""
    if isinstance(comm_stack, str):
        return {"s3": "file"}.get(comm_stack, comm_stack)
    get = comm_stack.get if hasattr(comm_stack, "get") else (lambda k, d=None: getattr(comm_stack, k, d))
    for key, kind in (("nvl", "nvl"), ("s3", "file"), ("shm", "shm"), ("ray", "inline")):
        if get(key, False):
            return kind
    return "inline"


def _as_arrays(data: Any, layout: FlatLayout | None) -> list[np.ndarray]:
    if torch.is_tensor(data):
        if layout is None:
            raise ValueError("a FlatLayout is needed to split a flat payload into per-tensor arrays")
        n_planes = max(1, data.numel() // layout.total)
        flat = data.detach().reshape(n_planes, -1)
        return [a for i in range(n_planes) for a in layout.to_ndarrays(flat[i])]
    return [np.asarray(a) for a in data]


def replace_remote_with_parameters_in_recordset(handle: ParamHandle, comm_stack: Any, *, endpoint_id: str,
                                                layout: FlatLayout | None = None, root: str | Path | None = None,
                                                folder_name: str = "comm_stack", file_name: str = "parameters",
                                                num_attempts: int = 3, store: Any = None, bucket: str | None = None) -> ParamHandle:
    """SENDER: move an inline payload onto the side channel and return the locator handle. ``store`` (an object store) +
    ``comm_stack="s3"`` / ``"link"``: the payload becomes the object ``{folder_name}/{endpoint_id}/{file_name}.npz`` of the bucket /
    of the fleet link's object service."""
    if store is not None and comm_stack == "link" and handle.kind == "inline":
        import tempfile

        key = f"{folder_name}/{endpoint_id}/{file_name}.npz"
        with tempfile.TemporaryDirectory() as td:
            store.upload(key, dump_model_parameters_to_file(Path(td) / "p.npz", _as_arrays(handle.data, layout)))
        return ParamHandle("link", key, {"endpoint_id": endpoint_id})
    if store is not None and comm_stack == "s3" and handle.kind == "inline":
        import tempfile

        key = f"{folder_name}/{endpoint_id}/{file_name}.npz"
        with tempfile.TemporaryDirectory() as td:
            store.upload(key, dump_model_parameters_to_file(Path(td) / "p.npz", _as_arrays(handle.data, layout)))
        return ParamHandle("s3", key, {"bucket": bucket, "endpoint_id": endpoint_id})
    kind = _comm_kind(comm_stack)
    if handle.kind != "inline" or kind in ("inline", "nvl"):
        return handle if kind != "nvl" else ParamHandle("nvl", None, dict(handle.meta))
    arrays = _as_arrays(handle.data, layout)
    if kind == "shm":
        shm, meta = set_parameters_shm(endpoint_id, arrays)
        _LIVE_SEGMENTS[endpoint_id] = shm
        return ParamHandle("shm", endpoint_id, {"metadata": meta.to_literal()})
    if root is None:
        raise ValueError("comm_stack.s3 needs a root directory for its objects")
    path = Path(root) / folder_name / endpoint_id / f"{file_name}.npz"
    dump_model_parameters_to_file(path, arrays)
    for attempt in range(max(1, num_attempts)):  # the reference polls S3 until the key is listable
        if path.exists():
            break
        time.sleep(0.05 * (attempt + 1))
    else:
        raise FileNotFoundError(f"object {path} did not become visible")
    return ParamHandle("file", str(path), {"endpoint_id": endpoint_id, "folder_name": folder_name, "file_name": file_name})


def replace_parameters_in_recordset_with_remote(handle: ParamHandle, *, layout: FlatLayout | None = None,
                                                as_flat: bool = False, device: torch.device | str = "cpu",
                                                num_attempts: int = 3) -> ParamHandle:
    """RECEIVER: resolve a locator back into an inline payload (fp32 ndarray list, or a flat tensor on
    ``device`` when ``as_flat``); ``nvl`` / ``inline`` handles pass through."""
    if handle.kind in ("inline", "nvl"):
        return handle
    if handle.kind == "shm":
        meta = ModelParametersMetadata.from_literal(handle.meta["metadata"])
        shm, views = get_parameters_shm(str(handle.data), meta, copy=True)
        shm.close()
        arrays = [v.astype(np.float32, copy=False) for v in views]
    elif handle.kind == "link":
        import tempfile

        if _LINK_STORE is None:
            raise RuntimeError(f"parameters were parked in the fleet link's object service ({handle.data}) but this process has no link")
        with tempfile.TemporaryDirectory() as td:
            arrays = load_model_parameters_from_file(_LINK_STORE.download(str(handle.data), Path(td) / "p.npz"))
    elif handle.kind == "s3":
        import tempfile

        from photon_b200.utils.objstore import remote_store_from_cfg

        store = remote_store_from_cfg({"s3_comm_config": {"bucket_name": handle.meta.get("bucket") or "checkpoints", "num_attempts": num_attempts}})
        if store is None:
            raise RuntimeError(f"parameters were parked in an object store ({handle.data}) but this process has no S3_ENDPOINT_URL / AWS_* settings")
        with tempfile.TemporaryDirectory() as td:
            arrays = load_model_parameters_from_file(store.download(str(handle.data), Path(td) / "p.npz"))
    elif handle.kind == "file":
        last: Exception | None = None
        for attempt in range(max(1, num_attempts)):
            try:
                arrays = load_model_parameters_from_file(str(handle.data))
                break
            except (OSError, ValueError) as e:
                last = e
                time.sleep(0.05 * (attempt + 1))
        else:
            raise FileNotFoundError(f"could not fetch {handle.data}: {last}")
    else:
        raise ValueError(f"unknown parameter handle kind {handle.kind!r}")
    if not as_flat:
        return ParamHandle("inline", arrays, dict(handle.meta))
    if layout is None:
        raise ValueError("as_flat needs a FlatLayout")
    n = len(layout.names)
    planes = []
    for i in range(0, len(arrays), n):
        flat = torch.zeros(layout.total, dtype=torch.float32)
        layout.from_ndarrays(flat, arrays[i:i + n])
        planes.append(flat)
    return ParamHandle("inline", torch.cat(planes).to(device), dict(handle.meta))


def release_remote_parameters(handle: ParamHandle) -> None:
    """Sender-side GC of a locator once every receiver has acknowledged it (the job of the reference's Ray
    ``custom_ray_garbage_collector`` thread / shm unlink / S3 delete; ref: photon/utils.py:73-144)."""
    if handle.kind == "shm":
        shm = _LIVE_SEGMENTS.pop(str(handle.data), None)
        if shm is not None:
            shm.close()
        unlink_quietly(str(handle.data))
    elif handle.kind == "file":
        Path(str(handle.data)).unlink(missing_ok=True)
    elif handle.kind == "link":
        if _LINK_STORE is not None:
            _LINK_STORE.delete(str(handle.data))
    elif handle.kind == "s3":
        from photon_b200.utils.objstore import remote_store_from_cfg

        store = remote_store_from_cfg({"s3_comm_config": {"bucket_name": handle.meta.get("bucket") or "checkpoints"}})
        if store is not None:
            store.delete(str(handle.data))


# ------------------------------------------------------------------ checkpoint API, reference-style free functions
def _store(cfg: Any) -> CheckpointStore:
    import os

    root = cfg["photon"].get("saving_path") or os.environ.get("PHOTON_SAVE_PATH", ".")
    from photon_b200.utils.objstore import remote_store_from_cfg

    return CheckpointStore(root, str(cfg["s3_comm_config"]["bucket_name"]), remote=remote_store_from_cfg(cfg))


def interpret_resume_round(cfg: Any, state_keys: Sequence[str]) -> int | None:
    return _store(cfg).interpret_resume_round(str(cfg["run_uuid"]), cfg["photon"].get("resume_round"), state_keys)


def import_checkpoints(cfg: Any, state_keys: Sequence[str]) -> int | None:
    return _store(cfg).import_checkpoints(cfg, state_keys)


def cleanup_checkpoints(cfg: Any, per_round: bool = False) -> None:
    _store(cfg).cleanup_checkpoints(str(cfg["run_uuid"]), per_round=per_round)


# ------------------------------------------------------------------ the reference's URI-based helpers (ref: s3_utils.py:75-212,1118-1641)
# They take ``s3://bucket/prefix`` strings (or plain directory paths) instead of a config: resolved against the configured S3 endpoint
# when there is one, else against the directory ``$PHOTON_SAVE_PATH/<bucket>``.
class NoCheckpointsFoundError(FileNotFoundError):
    """Nothing to resume from under the path that was looked up (ref: s3_utils.py:75-76)."""


def _uri_store(uri: str, leaf: bool = False) -> tuple[Any, str]:
    """(object store, key) of ``s3://bucket/key`` or of a local path. ``leaf``: the path names ONE object (the store is its
    directory); otherwise it names a prefix / directory (the store is rooted there, the key is empty)."""
    import os

    from photon_b200.utils.objstore import DirObjectStore, remote_store_from_cfg

    if str(uri).startswith("s3://"):
        bucket, _, prefix = str(uri)[len("s3://"):].partition("/")
        store = remote_store_from_cfg({"s3_comm_config": {"bucket_name": bucket}})
        if store is None:
            store = DirObjectStore(Path(os.environ.get("PHOTON_SAVE_PATH", ".")) / bucket, create=False)
        return store, prefix.strip("/")
    p = Path(uri)
    return (DirObjectStore(p.parent, create=False), p.name) if leaf else (DirObjectStore(p, create=False), "")


def list_objects(run_uuid_path: str) -> tuple[bool, list[str]]:
    """(anything there?, keys below the path, relative to it) (ref: s3_utils.py:114-155)."""
    store, prefix = _uri_store(run_uuid_path)
    keys = store.list(prefix + "/" if prefix else "")
    return bool(keys), [k[len(prefix) + 1:] if prefix else k for k in keys]


def delete_object(object_path: str) -> None:
    """(ref: s3_utils.py:79-111)"""
    store, key = _uri_store(object_path, leaf=True)
    store.delete(key)


def delete_remote_object(object_name: str) -> None:
    """Delete an object or everything below a prefix; a missing object is not an error (ref: s3_utils.py:1446-1475)."""
    store, key = _uri_store(object_name, leaf=True)
    store.delete(key)
    store.delete_prefix(key + "/")


def get_num_batches_from_checkpoint_name(checkpoint_name: str) -> int:
    """``ep3-ba1280-rank0.pt`` → 1280 (ref: s3_utils.py:1234-1258)."""
    import re

    m = re.search(r"-ba(\d+)-", checkpoint_name)
    if not m:
        raise ValueError(f"not a trainer checkpoint name (…-ba<batches>-…): {checkpoint_name}")
    return int(m.group(1))


def _rounds_under(store: Any, prefix: str, state_keys: Sequence[str]) -> list[int]:
    have: dict[int, set[str]] = {}
    base = f"{prefix}/server/" if prefix else "server/"
    for k in store.list(base):
        parts = k[len(base):].split("/")
        if len(parts) == 2 and parts[0].isdigit():
            have.setdefault(int(parts[0]), set()).add(parts[1])
    want = {"state.bin", *(f"{k}.npz" for k in state_keys)}
    return sorted(r for r, files in have.items() if want <= files)


def obtain_sorted_runs(run_uuid_path: str, state_keys: Sequence[str]) -> list[int]:
    """Round numbers with a COMPLETE server checkpoint under ``{run_uuid_path}/server/`` (ref: s3_utils.py:1261-1318; the reference
    names them "runs")."""
    store, prefix = _uri_store(run_uuid_path)
    rounds = _rounds_under(store, prefix, state_keys)
    if not rounds and not store.list(prefix + "/" if prefix else ""):
        raise NoCheckpointsFoundError(f"nothing under {run_uuid_path}")
    return rounds


def delete_rounds(run_uuid_path: str, state_keys: Sequence[str], end_idx: int | None = -1) -> None:
    """Delete the complete rounds ``[:end_idx]`` — the default keeps the newest (ref: s3_utils.py:1383-1443)."""
    store, prefix = _uri_store(run_uuid_path)
    base = f"{prefix}/server/" if prefix else "server/"
    for r in _rounds_under(store, prefix, state_keys)[:end_idx]:
        store.delete_prefix(f"{base}{r}/")


def delete_clients_checkpoints(run_uuid_path: str, end_idx: int | None = -1) -> None:
    """Per client folder, delete the trainer checkpoints ``[:end_idx]`` in batch order — the default keeps each client's newest
    (ref: s3_utils.py:1321-1380)."""
    store, prefix = _uri_store(run_uuid_path)
    per: dict[str, list[str]] = {}
    for k in store.list(prefix + "/" if prefix else ""):
        rel = k[len(prefix) + 1:] if prefix else k
        parts = rel.split("/")
        if len(parts) == 2 and parts[0].startswith("client_") and parts[1].startswith("ep") and parts[1].endswith(".pt"):
            per.setdefault(parts[0], []).append(k)
    for keys in per.values():
        for k in sorted(keys, key=lambda x: get_num_batches_from_checkpoint_name(x.rsplit("/", 1)[1]))[:end_idx]:
            store.delete(k)


def get_file_from_path(input_file_path: str, run_uuid: str = "", s3_comm_config: Any = None, tmp_dir: Any = None) -> Path:
    """A local path for ``input_file_path``: itself when it is a file, a download into ``tmp_dir`` when it is ``s3://…``
    (ref: s3_utils.py:1118-1189)."""
    del run_uuid, s3_comm_config
    if not str(input_file_path).startswith("s3://"):
        p = Path(input_file_path)
        if not p.is_file():
            raise FileNotFoundError(str(p))
        return p
    store, key = _uri_store(input_file_path, leaf=True)
    import os
    import tempfile

    if tmp_dir is None:
        d = Path(tempfile.mkdtemp(prefix="photon_dl_"))
    elif isinstance(tmp_dir, (str, os.PathLike)):
        d = Path(tmp_dir)
    else:
        d = Path(tmp_dir.name)          # a tempfile.TemporaryDirectory, like the reference passes
    return store.download(key, d / Path(key).name)


def extract_s3_comm_config_from_configrecord(s3_comm_config: Any) -> tuple[str, str, str]:
    """(endpoint_id, file_name, folder_name) of a side-channel locator: a ``ParamHandle`` or the reference's record / dict with
    ``endpoint_id`` / ``file_name`` / ``current_round`` (ref: s3_utils.py:158-212)."""
    meta = s3_comm_config.meta if isinstance(s3_comm_config, ParamHandle) else dict(s3_comm_config)
    folder = meta.get("folder_name", meta.get("current_round", "comm_stack"))
    return str(meta["endpoint_id"]), str(meta.get("file_name", "parameters")), str(folder)


def upload_server_checkpoint(cfg: Any, server_round: int, *, layout: FlatLayout, tensors: dict[str, torch.Tensor], state: dict[str, Any],
                             background: bool = False) -> Path:
    """One round of server state into the run's store (ref: s3_utils.py:480-548): ``state.bin`` + one npz per strategy state key."""
    return _store(cfg).upload_server_checkpoint(str(cfg["run_uuid"]), server_round, layout=layout, tensors=tensors, state=state, background=background)


def download_server_checkpoint(cfg: Any, server_round: int | None = None, *, layout: FlatLayout, state_keys: Sequence[str]
                               ) -> tuple[dict[str, torch.Tensor], dict[str, Any]]:
    """The round to resume from (``photon.resume_round`` when ``server_round`` is None) as (flat planes, state dict)
    (ref: s3_utils.py:551-727)."""
    st = _store(cfg)
    rnd = server_round if server_round is not None else st.interpret_resume_round(str(cfg["run_uuid"]), cfg["photon"].get("resume_round"), state_keys)
    if rnd is None:
        raise NoCheckpointsFoundError(f"run '{cfg['run_uuid']}' has no complete server checkpoint")
    return st.download_server_checkpoint(str(cfg["run_uuid"]), int(rnd), layout=layout, state_keys=state_keys)


def copy_old_checkpoints_to_new_run(cfg: Any, old_uuid: str, new_uuid: str, server_round: int, *, state_keys: Sequence[str],
                                    copy_client_checkpoints: bool = True, client_ids: Sequence[int] = ()) -> None:
    """(ref: s3_utils.py:1478-1641; includes the second momentum the reference forgets)"""
    _store(cfg).copy_old_checkpoints_to_new_run(old_uuid, new_uuid, server_round, state_keys=state_keys,
                                                copy_client_checkpoints=copy_client_checkpoints, client_ids=client_ids)

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
    ``comm_stack="s3"``: the payload becomes the object ``{folder_name}/{endpoint_id}/{file_name}.npz`` of the bucket."""
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

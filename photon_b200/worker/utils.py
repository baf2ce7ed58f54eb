"""Worker-side helpers: env patching for torch.distributed, result message, free-port pick
(ref: photon/worker/utils.py:47-159)."""
from __future__ import annotations

import contextlib
import os
import socket
from dataclasses import dataclass
from typing import Iterator


@dataclass
class WorkerResultMessage:
    """Posted on the node manager's result queue. ``n_samples == -1`` signals failure
    (ref: worker/utils.py:47-53, worker.py:427-448)."""

    n_samples: int
    delta: float          # seconds spent in the task
    worker_uuid: str
    cid: int | None = None
    error: str | None = None


def get_free_tcp_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return int(s.getsockname()[1])


@contextlib.contextmanager
def env_patcher(worker_rank: int, n_workers: int, master_port: int, devices: list[int] | None) -> Iterator[None]:
    """torchrun-style environment for the collaborating workers of ONE node: all workers of the
    node form one process group on 127.0.0.1 (ref: worker/utils.py:94-120)."""
    patch = {"RANK": str(worker_rank), "LOCAL_RANK": str(worker_rank), "WORLD_SIZE": str(n_workers),
             "LOCAL_WORLD_SIZE": str(n_workers), "NODE_RANK": "0", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(master_port),
             "TORCH_NCCL_ASYNC_ERROR_HANDLING": "1"}
    if devices:
        patch["APPOINTED_CUDA_DEVICE"] = ",".join(str(d) for d in devices)
    old = {k: os.environ.get(k) for k in patch}
    os.environ.update(patch)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@contextlib.contextmanager
def get_env_patcher(run_uuid: str, rank: str | int, master_port: str | int, world_size: str | int = 1, devices: list[int] | None = None) -> Iterator[None]:
    """The reference's name and argument order (ref: worker/utils.py:57-120): the environment of ONE worker of a node for the
    duration of a task, restored afterwards so the next task can form a fresh process group. ``run_uuid`` names the run in
    ``RUN_NAME`` (Composer read it from there)."""
    old = os.environ.get("RUN_NAME")
    os.environ["RUN_NAME"] = str(run_uuid)
    try:
        with env_patcher(int(rank), int(world_size), int(master_port), devices):
            yield
    finally:
        if old is None:
            os.environ.pop("RUN_NAME", None)
        else:
            os.environ["RUN_NAME"] = old

"""Worker process: one per device of a node; all workers of the node collaborate (DDP) on ONE
client at a time (ref: photon/worker/worker.py:207-520).

Loop: ``for task in iter(task_queue.get, None)`` with tasks ``(cid, "fit"|"evaluate")``.
The round's parameters come from the node manager's shm segment (zero-copy views), the per-client
config from the config segment; rank 0 writes the results (params / n_samples / metrics or eval loss)
to its own segments and posts a :class:`WorkerResultMessage`.  The Trainer is kept alive across
tasks (``external_trainer``).  On an exception the worker posts ``n_samples = -1`` and terminates —
the node manager re-queues the client and respawns the pool.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import queue
import time
import traceback
import uuid
from typing import Any

import numpy as np

from photon_b200.shm import constants as C
from photon_b200.shm.utils import (ModelParametersMetadata, get_dict_shm, get_parameters_shm, set_dict_shm,
                                   set_eval_loss_shm, set_n_samples_shm, set_parameters_shm, shm_exists)
from photon_b200.worker.utils import WorkerResultMessage, env_patcher


class Worker(mp.get_context("spawn").Process):  # type: ignore[misc,name-defined]
    def __init__(self, cfg_dict: dict[str, Any], nm_uuid: str, worker_rank: int, n_workers: int, task_queue: Any,
                 result_queue: Any, devices: list[int] | None = None) -> None:
        super().__init__(daemon=True, name=f"pb200-worker-{worker_rank}")
        self.cfg_dict, self.nm_uuid = cfg_dict, nm_uuid
        self.worker_rank, self.n_workers = worker_rank, n_workers
        self.task_queue, self.result_queue = task_queue, result_queue
        self.devices = devices
        self.worker_uuid = f"{nm_uuid}-w{worker_rank}-{uuid.uuid4().hex[:6]}"
        self.external_trainer: Any = None
        self._keep: list[Any] = []

    # ------------------------------------------------------------------ task bodies
    def _round_params(self) -> list[np.ndarray]:
        name = self.nm_uuid + C.NM_PARAMS_SHM
        t0 = time.time()
        while not shm_exists(name):  # the broadcast may still be landing (ref: worker.py:237-251)
            if time.time() - t0 > 120:
                raise TimeoutError(f"parameter segment {name} never appeared")
            time.sleep(0.05)
        meta = ModelParametersMetadata.from_literal(get_dict_shm(name + "_meta"))
        shm, views = get_parameters_shm(name, meta)
        self._keep.append(shm)
        return views

    def fit_action(self, cid: int, task_cfg: dict[str, Any]) -> tuple[int, dict[str, Any]]:
        from photon_b200.clients.llm_client_functions import llm_fit
        from photon_b200.config.composer import ConfigNode

        cfg = ConfigNode(self.cfg_dict)
        arrays, n, metrics, self.external_trainer = llm_fit(self.external_trainer, self._round_params(), task_cfg["fit_config"], cfg, cid,
                                                            as_ndarrays=True)
        if self.worker_rank == 0:
            shm, meta = set_parameters_shm(self.worker_uuid + C.W_PARAMS_SHM, arrays)
            self._keep.append(shm)
            set_dict_shm(self.worker_uuid + C.W_PARAMS_SHM + "_meta", meta.to_literal())
            set_n_samples_shm(self.worker_uuid, n)
            set_dict_shm(self.worker_uuid + C.W_METRICS_SHM, metrics)
        return n, metrics

    def evaluate_action(self, cid: int, task_cfg: dict[str, Any]) -> tuple[int, dict[str, Any]]:
        from photon_b200.clients.llm_client_functions import llm_eval
        from photon_b200.config.composer import ConfigNode

        cfg = ConfigNode(self.cfg_dict)
        loss, n, metrics, self.external_trainer = llm_eval(self.external_trainer, self._round_params(), task_cfg["eval_config"], cfg, cid)
        if self.worker_rank == 0:
            set_eval_loss_shm(self.worker_uuid, loss)
            set_n_samples_shm(self.worker_uuid, n)
            set_dict_shm(self.worker_uuid + C.W_METRICS_SHM, metrics)
        return n, metrics

    def process_task(self, cid: int, kind: str) -> None:
        t0 = time.time()
        all_cfg = get_dict_shm(self.nm_uuid + C.NM_CONFIG_SHM)
        task_cfg = all_cfg[str(cid)]
        if task_cfg.get("inject_failure") and self.worker_rank == 0:
            raise RuntimeError(f"fault injection: worker failure on client {cid}")
        if task_cfg.get("inject_hang") and self.worker_rank == 0:     # fault injection: alive, silent, never finishing
            time.sleep(3600)
        with env_patcher(self.worker_rank, self.n_workers, int(task_cfg["MASTER_PORT"]), self.devices):
            n, _ = (self.fit_action if kind == "fit" else self.evaluate_action)(cid, task_cfg)
        for h in self._keep[:-2]:
            h.close()
        self._keep = self._keep[-2:]
        self.result_queue.put(WorkerResultMessage(n_samples=int(n), delta=time.time() - t0, worker_uuid=self.worker_uuid, cid=cid))

    def _sweep_orphaned_segments(self) -> None:
        from photon_b200.shm.utils import close_all_shms, unlink_quietly

        close_all_shms(self.worker_uuid)
        unlink_quietly(self.worker_uuid + C.W_PARAMS_SHM + "_meta")
        if self.worker_rank == 0:
            close_all_shms(self.nm_uuid)
            unlink_quietly(self.nm_uuid + C.NM_PARAMS_SHM + "_meta")

    # ------------------------------------------------------------------------- loop
    def run(self) -> None:
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 2) // max(1, self.n_workers))))
        parent = os.getppid()
        while True:
            try:
                task = self.task_queue.get(timeout=2.0)
            except queue.Empty:
                if os.getppid() != parent:   # the node manager is gone (killed server): do not linger as an orphan,
                    self._sweep_orphaned_segments()   # and do not leave its /dev/shm segments (nor ours) behind
                    return
                continue
            if task is None:
                return
            cid, kind = task
            try:
                self.process_task(int(cid), kind)
            except BaseException as e:  # noqa: BLE001 - report, then die so the pool is rebuilt clean
                self.result_queue.put(WorkerResultMessage(n_samples=-1, delta=0.0, worker_uuid=self.worker_uuid, cid=int(cid),
                                                          error="".join(traceback.format_exception_only(type(e), e)).strip()))
                return


# ----------------------------------------------------------------------------- reference-named module functions (ref: worker.py:534-680)
def create_new_worker(config: Any, task_queue: Any, result_queue: Any, node_manager_uuid: str, run_uuid: str = "", parameters_metadata: Any = None,
                      worker_rank: int = 0, n_workers: int = 1, devices: list[int] | None = None) -> Worker:
    """A worker process object for rank ``worker_rank`` of a node (not started). ``config`` is the composed config (node or plain
    dict); the parameter metadata travels through the ``…_meta`` segment, so the argument is accepted and unused."""
    del run_uuid, parameters_metadata
    cfg_dict = config.to_container() if hasattr(config, "to_container") else dict(config)
    return Worker(cfg_dict, node_manager_uuid, int(worker_rank), int(n_workers), task_queue, result_queue, devices)


def start_worker(worker: Worker) -> None:
    worker.start()
    print(f"[node manager] worker {worker.worker_uuid} (rank {worker.worker_rank}) started", flush=True)


def get_training_results_from_worker(worker: Worker | None) -> tuple[list[np.ndarray], dict[str, Any], int] | None:
    """(parameters, metrics, n_samples) a worker left in its segments after a fit task; the arrays are copies, so the worker may
    reuse its segments (the reference returns the mapped segments too and closes them later)."""
    if worker is None:
        return None
    from photon_b200.shm.utils import get_n_samples_shm

    wu = worker.worker_uuid
    meta = ModelParametersMetadata.from_literal(get_dict_shm(wu + C.W_PARAMS_SHM + "_meta"))
    shm, views = get_parameters_shm(wu + C.W_PARAMS_SHM, meta, copy=True)
    shm.close()
    return views, dict(get_dict_shm(wu + C.W_METRICS_SHM)), get_n_samples_shm(wu)

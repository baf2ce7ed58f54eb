"""Server checkpoint store: layout, upload/download, resume-round resolution, restore
from another run, cleanup (the role of ref photon/server/s3_utils.py:215-727,1261-1641).

"Bucket" = a directory (``{saving_path}/{bucket_name}``) that is the working copy; when an S3 endpoint is configured
(``S3_ENDPOINT_URL`` / ``s3_comm_config.backend_kwargs.endpoint_url`` + ``AWS_*`` credentials → :mod:`photon_b200.utils.objstore`)
every round is also uploaded under the same keys, rounds that exist only remotely are found and fetched on resume, clean-up
deletes both copies, and client (trainer) checkpoints are mirrored by :class:`ClientCheckpointMirror`. The key layout is the reference's::

    {bucket}/{run_uuid}/server/{round}/state.bin
    {bucket}/{run_uuid}/server/{round}/current_server_parameters.npz      arr_i, sorted-name order
    {bucket}/{run_uuid}/server/{round}/current_momentum_vector.npz        (Nesterov/Mom/Adam/Yogi)
    {bucket}/{run_uuid}/server/{round}/current_second_momentum_vector.npz (Adam/Yogi)
    {bucket}/{run_uuid}/server/comm_stack/{endpoint}/parameters.npz       (s3 comm stack)
    {bucket}/{run_uuid}/client_{cid}/ep{E}-ba{B}-rank{R}.pt               (client checkpoints)

Deliberate fixes of reference quirks: the second momentum is restored
from its own file (the ref loads ``state["momentum"]`` into both) and is copied by
``copy_old_checkpoints_to_new_run``; cleanup honours ``bucket_name``.
"""
from __future__ import annotations

import pickle
import shutil
from pathlib import Path
from typing import Any, Sequence

import numpy as np
import torch

from photon_b200.utils.core import dump_model_parameters_to_file, load_model_parameters_from_file
from photon_b200.utils.flat import FlatLayout

STATE_FILE = "state.bin"


class CheckpointStore:
    def __init__(self, root: str | Path, bucket_name: str = "checkpoints", remote: Any = None) -> None:
        self.bucket = Path(root) / bucket_name
        self.bucket.mkdir(parents=True, exist_ok=True)
        self.remote = remote        # photon_b200.utils.objstore.ObjectStore or None

    def _key(self, path: Path) -> str:
        return path.relative_to(self.bucket).as_posix()

    def _push(self, path: Path) -> None:
        if self.remote is not None:
            self.remote.upload(self._key(path), path)

    def _ensure_local(self, path: Path) -> Path:
        """Fetch ``path`` from the object store when this host does not have it (resume on another machine)."""
        if not path.exists() and self.remote is not None and self.remote.exists(self._key(path)):
            self.remote.download(self._key(path), path)
        return path

    # -- key helpers ------------------------------------------------------------------
    def server_dir(self, run_uuid: str) -> Path:
        return self.bucket / run_uuid / "server"

    def round_dir(self, run_uuid: str, server_round: int) -> Path:
        return self.server_dir(run_uuid) / str(int(server_round))

    def client_dir(self, run_uuid: str, cid: int | str) -> Path:
        return self.bucket / run_uuid / f"client_{cid}"

    def list_objects(self, prefix: str | Path = "") -> list[str]:
        base = self.bucket / prefix
        local = {str(p.relative_to(self.bucket)) for p in base.rglob("*") if p.is_file()} if base.exists() else set()
        if self.remote is not None:
            pre = Path(prefix).as_posix() if str(prefix) else ""
            local |= set(self.remote.list(pre + "/" if pre else ""))
        return sorted(local)

    def delete_object(self, key: str | Path) -> None:
        p = self.bucket / key
        if p.is_dir():
            shutil.rmtree(p, ignore_errors=True)
        elif p.exists():
            p.unlink()
        if self.remote is not None:
            k = Path(key).as_posix()
            self.remote.delete(k)
            self.remote.delete_prefix(k + "/")

    # -- upload -----------------------------------------------------------------------
    def upload_server_checkpoint(self, run_uuid: str, server_round: int, *, layout: FlatLayout,
                                 tensors: dict[str, torch.Tensor], state: dict[str, Any], background: bool = False) -> Path:
        """Write one round: state.bin + one npz per strategy state key (ref: s3_utils.py:480-548).

        ``background=True``: the device → host copy of the planes and the pickling of ``state`` (whose history keeps growing)
        happen NOW, so the snapshot is the round's; the file writes (0.5 GB of npz per plane for MPT-125M, seconds for the larger
        models — the reference stalls its round loop on them, plus the S3 upload) run on a writer thread in submission order.
        :meth:`wait` blocks until everything submitted is on disk (and re-raises a writer error)."""
        d = self.round_dir(run_uuid, server_round)
        arrays = {key: layout.to_ndarrays(flat) for key, flat in tensors.items()}
        blob = pickle.dumps({"server_round": int(server_round), **state})

        def write() -> None:
            d.mkdir(parents=True, exist_ok=True)
            for key, arrs in arrays.items():
                dump_model_parameters_to_file(d / f"{key}.npz", arrs)
                self._push(d / f"{key}.npz")
            tmp = d / (STATE_FILE + ".tmp")
            tmp.write_bytes(blob)
            tmp.replace(d / STATE_FILE)  # state.bin last: its presence marks the round complete (locally and in the object store)
            self._push(d / STATE_FILE)

        if background:
            self.submit(write)
        else:
            write()
        return d

    # -- background writer ------------------------------------------------------------
    def submit(self, fn: Any) -> None:
        """Run ``fn()`` on the store's single writer thread, after everything submitted before it (at most 2 jobs queue up: a
        slow disk eventually back-pressures the round loop instead of piling snapshots up in host memory)."""
        import queue
        import threading

        if getattr(self, "_q", None) is None:
            self._q: Any = queue.Queue(maxsize=2)
            self._err: BaseException | None = None

            def loop() -> None:
                while True:
                    job = self._q.get()
                    try:
                        if job is None:
                            return
                        if self._err is None:
                            job()
                    except BaseException as e:  # noqa: BLE001 - surfaced by wait()
                        self._err = e
                    finally:
                        self._q.task_done()

            self._thread = threading.Thread(target=loop, name="photon-ckpt-writer", daemon=True)
            self._thread.start()
            import atexit

            atexit.register(self._drain_quietly)    # also when the run ends with an exception: finish what was snapshotted
        self._q.put(fn)

    def _drain_quietly(self) -> None:
        try:
            self.wait()
        except Exception as e:  # noqa: BLE001
            print(f"[checkpoint] {e}: {e.__cause__}", flush=True)

    def wait(self) -> None:
        """Block until every submitted write is on disk; raise what the writer thread raised, if anything."""
        if getattr(self, "_q", None) is not None:
            self._q.join()
            if self._err is not None:
                err, self._err = self._err, None
                raise RuntimeError("background checkpoint write failed") from err

    # -- discovery ----------------------------------------------------------------------
    def obtain_sorted_rounds(self, run_uuid: str, state_keys: Sequence[str]) -> list[int]:
        """Rounds that are COMPLETE: state.bin + every state-key file present (ref: s3_utils.py:1261-1318)."""
        base = self.server_dir(run_uuid)
        out = set()
        if base.exists():
            for p in base.iterdir():
                if p.is_dir() and p.name.isdigit() and (p / STATE_FILE).exists() and all((p / f"{k}.npz").exists() for k in state_keys):
                    out.add(int(p.name))
        if self.remote is not None:     # rounds another host uploaded
            have: dict[int, set[str]] = {}
            pre = f"{run_uuid}/server/"
            for k in self.remote.list(pre):
                parts = k[len(pre):].split("/")
                if len(parts) == 2 and parts[0].isdigit():
                    have.setdefault(int(parts[0]), set()).add(parts[1])
            want = {STATE_FILE, *(f"{k}.npz" for k in state_keys)}
            out |= {r for r, files in have.items() if want <= files}
        return sorted(out)

    def interpret_resume_round(self, run_uuid: str, resume_round: int | None, state_keys: Sequence[str]) -> int | None:
        """``None`` → start from scratch; ``-1`` → latest complete round (None if there is none);
        ``k >= 0`` → that round, which must be complete (ref: s3_utils.py:215-272)."""
        if resume_round is None:
            return None
        rounds = self.obtain_sorted_rounds(run_uuid, state_keys)
        if resume_round < 0:
            idx = len(rounds) + resume_round
            return rounds[idx] if 0 <= idx < len(rounds) else None
        if resume_round not in rounds:
            raise FileNotFoundError(f"round {resume_round} of run '{run_uuid}' is missing or incomplete (have {rounds})")
        return resume_round

    # -- download -----------------------------------------------------------------------
    def download_server_checkpoint(self, run_uuid: str, server_round: int, *, layout: FlatLayout,
                                   state_keys: Sequence[str]) -> tuple[dict[str, torch.Tensor], dict[str, Any]]:
        d = self.round_dir(run_uuid, server_round)
        state = load_server_state(self._ensure_local(d / STATE_FILE))
        tensors: dict[str, torch.Tensor] = {}
        for key in state_keys:
            arrays = load_model_parameters_from_file(self._ensure_local(d / f"{key}.npz"))
            flat = torch.zeros(layout.total, dtype=torch.float32)
            layout.from_ndarrays(flat, arrays)
            tensors[key] = flat
        return tensors, state

    # -- restore from another run ---------------------------------------------------------
    def copy_old_checkpoints_to_new_run(self, old_uuid: str, new_uuid: str, server_round: int, *, state_keys: Sequence[str],
                                        copy_client_checkpoints: bool = True, client_ids: Sequence[int] = ()) -> None:
        src, dst = self.round_dir(old_uuid, server_round), self.round_dir(new_uuid, server_round)
        dst.mkdir(parents=True, exist_ok=True)
        for key in state_keys:  # includes the second momentum (the reference forgets it)
            shutil.copy2(self._ensure_local(src / f"{key}.npz"), dst / f"{key}.npz")
            self._push(dst / f"{key}.npz")
        shutil.copy2(self._ensure_local(src / STATE_FILE), dst / STATE_FILE)
        self._push(dst / STATE_FILE)
        if copy_client_checkpoints:
            for cid in client_ids:
                s = self.client_dir(old_uuid, cid)
                if s.exists():
                    shutil.copytree(s, self.client_dir(new_uuid, cid), dirs_exist_ok=True)

    def import_checkpoints(self, cfg: Any, state_keys: Sequence[str]) -> int | None:
        """Seed THIS run (``cfg.run_uuid``) from ``photon.restore_run_uuid``: resolve the round to restore
        (``photon.resume_round``; -1 = newest complete), copy its server checkpoint (+ client checkpoints when
        ``photon.copy_client_checkpoints``) under the new run id; no-op if this run already has checkpoints
        (ref: server/s3_utils.py:275-345). Returns the restored round (also written back to the config)."""
        ph = cfg["photon"]
        src, dst = ph.get("restore_run_uuid"), cfg.get("run_uuid")
        if not src or not dst:
            raise ValueError("import_checkpoints needs both run_uuid and photon.restore_run_uuid")
        want = ph.get("resume_round", -1)
        rnd = self.interpret_resume_round(str(src), -1 if want is None else want, state_keys)   # restoring a run with no round named = its newest
        if rnd is None or self.obtain_sorted_rounds(str(dst), state_keys):
            return None
        self.copy_old_checkpoints_to_new_run(str(src), str(dst), rnd, state_keys=state_keys,
                                             copy_client_checkpoints=bool(ph.get("copy_client_checkpoints", True)),
                                             client_ids=range(int(cfg["fl"]["n_total_clients"])))
        # client checkpoints normally live under llm_config.save_folder, whose path carries the run id
        # (``…/{run_uuid}/clients/client_{cid}/``): bring the old run's over so optimizer state / data position survive
        sf = (cfg.get("llm_config") or {}).get("save_folder")
        if bool(ph.get("copy_client_checkpoints", True)) and sf and str(dst) in Path(str(sf)).parts:
            old_sf = Path(*[str(src) if part == str(dst) else part for part in Path(str(sf)).parts])   # whole path components only
            for cid in range(int(cfg["fl"]["n_total_clients"])):
                s_dir = old_sf / f"client_{cid}"
                if s_dir.is_dir():
                    shutil.copytree(s_dir, Path(str(sf)) / f"client_{cid}", dirs_exist_ok=True, symlinks=True)
        ph["resume_round"] = rnd
        return rnd

    def obtain_sorted_runs(self) -> list[str]:
        """Run ids in the bucket that hold at least one server round, oldest first by modification time
        (ref: server/s3_utils.py:1261-1318 sorts run folders to pick what to garbage-collect)."""
        runs = [p for p in self.bucket.iterdir() if p.is_dir() and (p / "server").is_dir()] if self.bucket.exists() else []
        return [p.name for p in sorted(runs, key=lambda p: p.stat().st_mtime)]

    # -- cleanup ------------------------------------------------------------------------
    def delete_rounds(self, run_uuid: str, keep_last: int = 0) -> None:
        base = self.server_dir(run_uuid)
        rounds = sorted(int(p.name) for p in base.iterdir() if p.is_dir() and p.name.isdigit()) if base.exists() else []
        complete = [r for r in rounds if (base / str(r) / STATE_FILE).exists()]
        remote_rounds: set[int] = set()
        if self.remote is not None:
            pre = f"{run_uuid}/server/"
            for k in self.remote.list(pre):
                head = k[len(pre):].split("/")[0]
                if head.isdigit():
                    remote_rounds.add(int(head))
                    if k.endswith("/" + STATE_FILE) and int(head) not in complete:
                        complete.append(int(head))
            complete.sort()
        keep = set(complete[-keep_last:]) if keep_last else set()
        for r in sorted(set(rounds) | remote_rounds):
            if r not in keep:
                shutil.rmtree(base / str(r), ignore_errors=True)
                if r in remote_rounds:
                    self.remote.delete_prefix(f"{run_uuid}/server/{r}/")

    def delete_clients_checkpoints(self, run_uuid: str, keep_latest: bool = False) -> None:
        base = self.bucket / run_uuid
        if not base.exists():
            return
        for cdir in base.glob("client_*"):
            files = sorted(cdir.glob("ep*-ba*-rank*.pt"), key=lambda p: int(p.name.split("-ba")[1].split("-")[0]))
            for f in (files[:-1] if keep_latest else files):
                f.unlink()

    def cleanup_checkpoints(self, run_uuid: str, per_round: bool = False) -> None:
        """``per_round`` keeps only the newest round / client checkpoint; otherwise remove everything."""
        if per_round:
            self.delete_rounds(run_uuid, keep_last=1)
            self.delete_clients_checkpoints(run_uuid, keep_latest=True)
        else:
            shutil.rmtree(self.bucket / run_uuid, ignore_errors=True)
            if self.remote is not None:
                self.remote.delete_prefix(f"{run_uuid}/")


class ClientCheckpointMirror:
    """Client (trainer) checkpoints through the object store, under the reference's keys ``{run_uuid}/client_{cid}/ep…-rank{r}.pt``
    (ref: Composer saves to ``save_folder: s3://…`` through its RemoteUploaderDownloader). With nodes on several machines a client is
    trained wherever there is a free node: its optimizer moments and data position follow it through the bucket instead of staying
    on the disk of the machine that trained it last. ``pull`` runs before the mid-round resume / skip decision (which looks at the
    local folder), ``push`` after the client's training."""

    def __init__(self, remote: Any, run_uuid: str, keep: int = 1) -> None:
        self.remote, self.run_uuid, self.keep = remote, str(run_uuid), max(1, int(keep))

    def _prefix(self, cid: int | str) -> str:
        return f"{self.run_uuid}/client_{cid}/"

    @staticmethod
    def _batches(name: str) -> int:
        import re

        m = re.search(r"-ba(\d+)-", name)
        return int(m.group(1)) if m else -1

    def pull(self, cid: int | str, folder: str | Path) -> list[str]:
        """Fetch the newest ``keep`` checkpoints (every rank file of them) that are not in ``folder`` yet."""
        names = [k.rsplit("/", 1)[1] for k in self.remote.list(self._prefix(cid)) if k.endswith(".pt")]
        newest = sorted({self._batches(n) for n in names if self._batches(n) >= 0})[-self.keep:]
        got = []
        for n in names:
            if self._batches(n) in newest and not (Path(folder) / n).exists():
                self.remote.download(self._prefix(cid) + n, Path(folder) / n)
                got.append(n)
        return got

    def push(self, cid: int | str, folder: str | Path) -> list[str]:
        """Upload the checkpoints of ``folder`` the bucket does not have; drop remote ones older than the newest ``keep``."""
        p = Path(folder)
        if not p.is_dir():
            return []
        have = {k.rsplit("/", 1)[1] for k in self.remote.list(self._prefix(cid))}
        sent = []
        for f in sorted(p.glob("ep*-ba*-rank*.pt")):
            if f.is_file() and not f.is_symlink() and f.name not in have:
                self.remote.upload(self._prefix(cid) + f.name, f)
                sent.append(f.name)
        names = have | set(sent)
        newest = sorted({self._batches(n) for n in names if self._batches(n) >= 0})[-self.keep:]
        for n in names:
            if n.endswith(".pt") and self._batches(n) not in newest:
                self.remote.delete(self._prefix(cid) + n)
        return sent


def client_checkpoint_mirror(cfg: Any) -> ClientCheckpointMirror | None:
    """The mirror of this run, or None when no object store is configured (client checkpoints then stay on the local disk)."""
    from photon_b200.utils.objstore import remote_store_from_cfg

    remote = remote_store_from_cfg(cfg)
    if remote is None:
        return None
    keep = int((cfg.get("llm_config") or {}).get("save_num_checkpoints_to_keep", 1) or 1)
    return ClientCheckpointMirror(remote, str(cfg["run_uuid"]), keep=keep if keep > 0 else 1)


class _Foreign:
    """Stand-in for a class this process cannot import while unpickling (attributes land in ``__dict__``)."""

    def __setstate__(self, state: Any) -> None:
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return type(name, (_Foreign,), {"__module__": module})


def load_server_state(path: str | Path) -> dict[str, Any]:
    """``state.bin`` of either framework. The reference pickles the same five fields (``server_round``, ``history``,
    ``time_offset``, ``client_state`` as a literal string, ``server_steps_cumulative``; ref: s3_utils.py:374-389) but its
    ``history`` is a Flower ``History`` subclass this process cannot import: it is read through a tolerant unpickler and its
    five metric stores are copied into our :class:`WandbHistory`, so a federation started with the reference resumes here."""
    from photon_b200.wandb_history import History, WandbHistory

    with open(path, "rb") as f:
        state = _TolerantUnpickler(f).load()  # noqa: S301 - a checkpoint the user points us at
    hist = state.get("history")
    if hist is not None and not isinstance(hist, History):
        mine = WandbHistory(False)
        for attr in ("losses_distributed", "losses_centralized", "metrics_distributed_fit", "metrics_distributed", "metrics_centralized"):
            val = getattr(hist, attr, None)
            if val is not None:
                setattr(mine, attr, type(getattr(mine, attr))(val))
        state["history"] = mine
    return state


def load_pretrained_model_from_path(path: str | Path) -> list[np.ndarray]:
    """npz / npzc / bin from a local path or ``s3://bucket/key`` (needs ``S3_ENDPOINT_URL`` + ``AWS_*`` credentials in the
    environment, like the reference's downloader) (ref: s3_utils.py:1192-1231)."""
    if str(path).startswith("s3://"):
        import tempfile

        from photon_b200.utils.objstore import remote_store_from_cfg

        bucket, _, key = str(path)[len("s3://"):].partition("/")
        store = remote_store_from_cfg({"s3_comm_config": {"bucket_name": bucket}})
        if store is None or not key:
            raise RuntimeError(f"{path}: set S3_ENDPOINT_URL and AWS_ACCESS_KEY_ID / AWS_SECRET_ACCESS_KEY to read from an object store, "
                               "or copy the file locally")
        with tempfile.TemporaryDirectory() as td:
            return load_model_parameters_from_file(store.download(key, Path(td) / Path(key).name))
    return load_model_parameters_from_file(path)

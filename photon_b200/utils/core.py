"""Core helpers shared by server and clients (the role of ref photon/utils.py):
parameter get/set by (filtered, sorted) name, freeze lists, npz/bin model files,
norm helpers, device/core counts, parameter sanity checker, wandb init."""
from __future__ import annotations

import os
import pickle
from pathlib import Path
from typing import Any, Sequence

import numpy as np
import torch

_STRIP = ("model.", "module.", "_fsdp_wrapped_module.", "_checkpoint_wrapped_module.")


def clean_parameter_name(name: str) -> str:
    """Strip wrapper prefixes so names match the bare ``transformer.*`` tree
    (ref: photon/utils.py:602-637)."""
    changed = True
    while changed:
        changed = False
        for s in _STRIP:
            if s in name:
                name = name.replace(s, "")
                changed = True
    return name


def get_list_of_parameters_names(model: torch.nn.Module, key_filter: str | None = None) -> list[str]:
    names = sorted(clean_parameter_name(n) for n, p in model.named_parameters() if p.requires_grad)
    return [n for n in names if key_filter is None or key_filter in n]


def construct_parameters_dict(names: Sequence[str], arrays: Sequence[Any], key_filter: str | None = None) -> dict[str, Any]:
    """Zip sorted names with payload arrays, keeping only keys that contain ``key_filter``
    (ref: photon/utils.py:640-670)."""
    if len(names) != len(arrays):
        raise ValueError(f"{len(names)} names vs {len(arrays)} arrays")
    return {n: a for n, a in zip(names, arrays) if key_filter is None or key_filter in n}


def parameters_checker(a: Sequence[np.ndarray], b: Sequence[np.ndarray], *, expect_equal: bool, what: str = "") -> None:
    """Shapes must match; values must be all-equal / not-all-equal (ref: photon/utils.py:147-224)."""
    if len(a) != len(b):
        raise AssertionError(f"{what}: {len(a)} vs {len(b)} tensors")
    same = True
    for i, (x, y) in enumerate(zip(a, b)):
        if x.shape != y.shape:
            raise AssertionError(f"{what}: tensor {i} shape {x.shape} vs {y.shape}")
        same = same and bool(np.array_equal(x, y))
    if expect_equal and not same:
        raise AssertionError(f"{what}: parameters differ but were expected to be equal")
    if not expect_equal and same:
        raise AssertionError(f"{what}: parameters are identical but were expected to change")


# ------------------------------------------------------------ parameter get / set on a live trainer
def _flat_of(trainer: Any) -> Any:
    st = getattr(trainer, "state", trainer)
    flat = getattr(st, "flat", None)
    if flat is None:
        raise TypeError("expected a photon_b200 Trainer (or its state) holding flat parameter storage")
    return flat


def get_trainable_params_dict(model_or_trainer: Any, *, sort_dict: bool = True,
                              no_detach_and_clone: bool = False) -> dict[str, torch.Tensor]:
    """name → tensor for every trainable parameter (ref: photon/utils.py:247-319).

    The reference needs an FSDP ``summon_full_params`` + CPU offload here; our masters live in one flat
    buffer that every rank holds in full, so this is a dictionary of views (``no_detach_and_clone``)
    or of clones, already in sorted-name order.  A plain ``nn.Module`` is accepted too."""
    if isinstance(model_or_trainer, torch.nn.Module):
        items = [(clean_parameter_name(n), p) for n, p in model_or_trainer.named_parameters() if p.requires_grad]
        if sort_dict:
            items.sort(key=lambda kv: kv[0])
        return {n: (p if no_detach_and_clone else p.detach().clone()) for n, p in items}
    flat = _flat_of(model_or_trainer)
    views = flat.layout.views(flat.full_params())      # assembled from the owners when the parameters are fully sharded
    return {n: (v if no_detach_and_clone else v.detach().clone()) for n, v in zip(flat.layout.names, views)}


def get_parameters_from_state(_config: Any, trainer: Any) -> list[np.ndarray]:
    """The model as the reference's payload: one fp32 ndarray per tensor, sorted by name (ref: :227-244)."""
    return _flat_of(trainer).to_ndarrays()


def set_trainer_trainable_params_dict(trainer: Any, params: dict[str, Any]) -> None:
    """Overwrite the named tensors (a subset is fine) and refresh the bf16 compute copy
    (ref: photon/utils.py:390-478 — there a pickled rank-0 broadcast under FSDP)."""
    flat = _flat_of(trainer)
    full = flat.full_params()
    with torch.no_grad():
        for name, value in params.items():
            view = flat.layout.view(full, clean_parameter_name(name))
            t = torch.as_tensor(value)
            if tuple(t.shape) != tuple(view.shape):
                raise ValueError(f"{name}: shape {tuple(t.shape)} != model {tuple(view.shape)}")
            view.copy_(t.to(view.device, view.dtype))
    if flat.is_sharded:
        flat.load_full_params(full)
    backend = getattr(getattr(trainer, "state", trainer), "backend", None)
    if backend is not None:
        backend.params_updated()


def set_trainer_params_from_ndarrays(arrays: Sequence[np.ndarray], trainer: Any, key_filter: str | None = None) -> None:
    """Install a sorted-name payload; with ``key_filter`` the arrays correspond to the names containing it
    (ref: photon/utils.py:481-540)."""
    flat = _flat_of(trainer)
    names = [n for n in flat.layout.names if key_filter is None or key_filter in n]
    arrays = list(arrays)
    try:
        set_trainer_trainable_params_dict(trainer, construct_parameters_dict(names, arrays))
    except ValueError as sorted_err:
        # the reference's fallback for payloads written in MODEL DEFINITION order instead of sorted-name order
        # (ref: photon/utils.py:515-540); only taken when the sorted interpretation does not even fit the shapes
        model = getattr(getattr(getattr(trainer, "state", trainer), "backend", None), "model", None)
        if model is None:
            raise
        unsorted = [clean_parameter_name(n) for n, p in model.named_parameters() if p.requires_grad]
        unsorted = [n for n in unsorted if key_filter is None or key_filter in n]
        if unsorted == names or len(unsorted) != len(arrays):
            raise
        try:
            set_trainer_trainable_params_dict(trainer, construct_parameters_dict(unsorted, arrays))
        except ValueError:
            raise sorted_err from None
        print("[params] payload did not fit the sorted-name layout; installed it in model-definition order")


def get_wte_parameters_from_trainer(trainer: Any) -> np.ndarray:
    """The (unique) token-embedding matrix (ref: photon/utils.py:543-582)."""
    flat = _flat_of(trainer)
    hits = [n for n in flat.layout.names if "wte" in n]
    if len(hits) != 1:
        raise ValueError("There are no WTE parameters" if not hits else "WTE parameters are not unique")
    return flat.layout.view(flat.full_params(), hits[0]).detach().cpu().numpy().copy()


def set_wte_parameters_to_trainer(trainer: Any, wte_parameters: np.ndarray) -> None:
    """Transplant only the token embedding (ref: photon/utils.py:585-599)."""
    flat = _flat_of(trainer)
    hits = [n for n in flat.layout.names if "wte" in n]
    if len(hits) != 1:
        raise ValueError("There are no WTE parameters" if not hits else "WTE parameters are not unique")
    set_trainer_trainable_params_dict(trainer, {hits[0]: np.asarray(wte_parameters, dtype=np.float32)})


# ------------------------------------------------------------------- model files (npz / npzc / bin)
def dump_model_parameters_to_file(path: str | os.PathLike, arrays: Sequence[np.ndarray], compressed: bool = False) -> Path:
    """``np.savez(*arrays)`` → keys ``arr_0..arr_{n-1}`` in sorted-name order (ref: photon/utils.py:733-761)."""
    p = Path(path)
    p.parent.mkdir(parents=True, exist_ok=True)
    tmp = p.with_name(p.name + ".tmp")
    with open(tmp, "wb") as f:
        (np.savez_compressed if compressed else np.savez)(f, *arrays)
    os.replace(tmp, p)
    return p


def load_model_parameters_from_file(path: str | os.PathLike) -> list[np.ndarray]:
    """Accepts ``.npz``, ``.npzc`` (compressed) and ``.bin`` (pickled list / Flower-like Parameters)."""
    p = Path(path)
    if p.suffix in (".npz", ".npzc"):
        with np.load(p, allow_pickle=False) as z:
            keys = sorted(z.files, key=lambda k: int(k.split("_")[1]))
            return [z[k] for k in keys]
    if p.suffix == ".bin":
        with open(p, "rb") as f:
            obj = pickle.load(f)  # noqa: S301 - our own checkpoints
        if hasattr(obj, "tensors"):
            import io

            return [np.load(io.BytesIO(t), allow_pickle=False) for t in obj.tensors]
        return [np.asarray(a) for a in obj]
    raise ValueError(f"unsupported parameter file {p}")


# ----------------------------------------------------------------------------------- norms
def sum_of_squares(arrays: Sequence[np.ndarray | torch.Tensor]) -> float:
    return float(sum(float((torch.as_tensor(a).double() ** 2).sum()) for a in arrays))


def l2_norm(arrays: Sequence[np.ndarray | torch.Tensor]) -> float:
    return float(np.sqrt(sum_of_squares(arrays)))


def l2_norm_of_momenta(optimizer_or_state: Any) -> tuple[float, float]:
    """(‖exp_avg‖, ‖exp_avg_sq‖) of the local optimizer (ref: photon/utils.py:878-908). Accepts our flat
    optimizer (``full_moments()``) or a torch-style ``{param: {"exp_avg", "exp_avg_sq"}}`` state dict."""
    if hasattr(optimizer_or_state, "full_moments"):
        m, v = optimizer_or_state.full_moments()
        return float(m.double().norm()), float(v.double().norm())
    vals = list(optimizer_or_state.values())
    return l2_norm([s["exp_avg"] for s in vals]), l2_norm([s["exp_avg_sq"] for s in vals])


def chunks_idx(list_of_stuff: Sequence[Any], n_chunks: int) -> Any:
    """(start, end) of ``n_chunks`` near-equal contiguous chunks, longer ones first (ref: photon/utils.py:819-841)."""
    d, r = divmod(len(list_of_stuff), n_chunks)
    lo = 0
    for i in range(n_chunks):
        hi = lo + d + (1 if i < r else 0)
        yield lo, hi
        lo = hi


def is_literal_for_ast(s: str) -> bool:
    """True when ``ast.literal_eval`` accepts the string (ref: photon/utils.py:1066-1084)."""
    import ast

    try:
        ast.literal_eval(s)
    except (ValueError, SyntaxError):
        return False
    return True


# ----------------------------------------------------------------------------- environment
def get_n_cuda_devices() -> int:
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def get_n_cpu_cores() -> int | None:
    conc = os.environ.get("CPU_CONCURRENCY")
    return int(conc) if conc else os.cpu_count()


def get_device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def appointed_cuda_devices() -> list[int]:
    """``APPOINTED_CUDA_DEVICE="0,1,…"`` contract of the launch scripts (ref: worker/utils.py:94-120)."""
    s = os.environ.get("APPOINTED_CUDA_DEVICE", "")
    return [int(x) for x in s.split(",") if x.strip() != ""]


def wandb_init(enabled: bool, **kwargs: Any) -> Any:
    """Returns a wandb run or None (no network here → offline / disabled; ref: photon/utils.py:780-816)."""
    if not enabled:
        return None
    try:
        import wandb  # type: ignore[import-not-found]

        kwargs.setdefault("mode", "offline")
        return wandb.init(**kwargs)
    except Exception as e:  # noqa: BLE001
        print(f"[wandb] unavailable ({e}); metrics stay in the History object")
        return None


# ------------------------------------------------------------------ freezing / unigram / object stores
def freeze_blocks(model: torch.nn.Module, frozen_layers: Sequence[str] | None = None,
                  unfrozen_layers: Sequence[str] | None = None) -> list[str]:
    """``requires_grad=False`` for every parameter matched by ``frozen_layers`` — or, when
    ``unfrozen_layers`` is given, for everything NOT matched by it (ref: photon/utils.py:322-387).
    Returns the frozen parameter names."""
    from photon_b200.train.backend import apply_freeze

    return apply_freeze(model, list(frozen_layers or []) or None, list(unfrozen_layers or []) or None)


def add_unigram_metrics(trainer: Any, unigram_freq: dict[int, int] | dict[str, int]) -> None:
    """Attach the four unigram-normalised metrics to a live trainer (train + every eval label)
    (ref: clients/trainer_utils.py:278-327)."""
    from photon_b200.metrics.language import build_metrics, unigram_log_probs

    st = trainer.state
    logp = unigram_log_probs({int(k): int(v) for k, v in unigram_freq.items()}, trainer.model_cfg.vocab_size)
    cur = getattr(st.backend, "unigram_log_probs", None)
    if cur is not None and cur.shape == logp.shape:
        cur.copy_(logp)   # in place: a captured CUDA graph keeps reading this buffer
    else:
        st.backend.unigram_log_probs = logp.to(st.flat.params.device)
    st.train_metrics = build_metrics(True)
    st.eval_metrics = {lbl: build_metrics(True) for lbl in st.eval_metrics}


class RemoteUploaderDownloader:
    """Object-store façade with the call surface the reference gets from Composer's ``RemoteUploaderDownloader``
    (ref: photon/utils.py:955-1014): ``upload_file`` / ``download_file`` / ``list_objects`` / ``delete_object`` with retries, over
    an :class:`photon_b200.utils.objstore.ObjectStore` — the S3 bucket when an endpoint + credentials are configured
    (``S3_ENDPOINT_URL`` / ``AWS_*``), a directory under ``root`` otherwise."""

    def __init__(self, root: str | os.PathLike, bucket_name: str, prefix: str = "", num_attempts: int = 3, store: Any = None) -> None:
        from photon_b200.utils.objstore import DirObjectStore, remote_store_from_cfg

        self.remote_bucket_name, self.backend_kwargs = bucket_name, {"prefix": prefix}
        self.num_attempts = max(1, int(num_attempts))
        if store is None:
            store = remote_store_from_cfg({"s3_comm_config": {"bucket_name": bucket_name, "num_attempts": num_attempts,
                                                              "backend_kwargs": {"prefix": prefix}}})
        self.store = store if store is not None else DirObjectStore(Path(root) / bucket_name / prefix)
        self.run_name: str | None = None

    def init(self, run_name: str | None = None) -> None:
        self.run_name = run_name

    def _check_workers(self) -> None:      # Composer's uploader has worker processes to check; nothing to do here
        return None

    def upload_file(self, remote_file_name: str, file_path: str | os.PathLike, overwrite: bool = True, state: Any = None) -> None:
        del state
        if not overwrite and self.store.exists(remote_file_name):
            raise FileExistsError(remote_file_name)
        self.store.upload(remote_file_name, file_path)      # readers never see partial objects (atomic rename / S3 PUT)

    def download_file(self, remote_file_name: str, destination: str | os.PathLike, overwrite: bool = True) -> None:
        if os.path.exists(destination) and not overwrite:
            raise FileExistsError(str(destination))
        self.store.download(remote_file_name, destination)

    def list_objects(self, prefix: str = "") -> list[str]:
        return [k for k in self.store.list(prefix) if not k.endswith(".part")]

    def delete_object(self, remote_file_name: str) -> None:
        self.store.delete(remote_file_name)

    def close(self) -> None:
        pass


def create_remote_up_down(bucket_name: str, prefix: str, run_uuid: str | None, num_attempts: int,
                          client_config: dict[str, Any] | None = None, *, root: str | os.PathLike | None = None,
                          **_unused: Any) -> RemoteUploaderDownloader:
    """(ref: photon/utils.py:955-1014) ``root`` (directory fall-back) defaults to ``$PHOTON_SAVE_PATH`` (or ./runs)."""
    del client_config
    r = RemoteUploaderDownloader(root or os.environ.get("PHOTON_SAVE_PATH", "runs"), bucket_name, prefix, num_attempts)
    r.init(run_name=run_uuid)
    return r


def upload_file_to_s3(remote_up_down: RemoteUploaderDownloader, remote_file_name: str, local_file_name: str | os.PathLike) -> None:
    """(ref: photon/utils.py:687-699)"""
    remote_up_down.upload_file(remote_file_name, local_file_name, overwrite=True)


def download_file_from_s3(remote_up_down: RemoteUploaderDownloader, remote_file_name: str, local_file_name: str | os.PathLike) -> None:
    """(ref: photon/utils.py:673-684)"""
    remote_up_down.download_file(remote_file_name, local_file_name, overwrite=True)


# ------------------------------------------------------------------------------------------- small reference-named helpers
class NoOpContextManager:
    """``with NoOpContextManager(): ...`` does nothing (ref: photon/utils.py:56-71)."""

    def __enter__(self) -> None:
        return None

    def __exit__(self, *exc: Any) -> None:
        return None


def merge_freq_dicts(a: dict[int, int], b: dict[int, int]) -> dict[int, int]:
    """Token-count maps of two streams added key by key (ref: photon/utils.py:1017-1036); the n-ary, JSON-keyed form used by the
    client config code is ``clients.llm_config_functions.merge_freq_dicts``."""
    out = dict(a)
    for k, v in b.items():
        out[k] = out.get(k, 0) + v
    return out


def get_unigram_probabilities_tensor(stream_freq_dict: dict[int, int]) -> torch.Tensor:
    """Dense ``p[token]`` up to the largest token id seen (ref: photon/utils.py:1039-1063). The metrics use
    ``metrics.language.unigram_log_probs`` (vocabulary-sized, one pseudo-count for unseen ids so the CE stays finite)."""
    ids = torch.tensor([int(k) for k in stream_freq_dict], dtype=torch.long)
    counts = torch.tensor([float(v) for v in stream_freq_dict.values()], dtype=torch.float64)
    p = torch.zeros(int(ids.max()) + 1, dtype=torch.float64)
    p[ids] = counts / counts.sum()
    return p.to(torch.float32)


def set_parameters(net: torch.nn.Module, parameters: Sequence[np.ndarray], device: str = "cpu") -> None:
    """Install a sorted-name payload into a plain module (ref: photon/utils.py:764-776): the arrays pair with the module's
    TRAINABLE parameters in sorted-name order; the module is left in eval mode like the reference's."""
    net.eval()
    names = list(get_trainable_params_dict(net))
    if len(names) != len(parameters):
        raise ValueError(f"{len(parameters)} arrays for {len(names)} trainable tensors")
    lookup = {clean_parameter_name(n): p for n, p in net.named_parameters()}
    with torch.no_grad():
        for n, a in zip(names, parameters):
            lookup[n].copy_(torch.as_tensor(a, device=device).to(lookup[n].dtype))


def custom_ray_garbage_collector(garbage_queue: Any, list_of_threads: list[Any] | None = None, timeout: float = 300.0, *,
                                 join_at_the_end: bool = True) -> Any:
    """Context manager: a background thread frees the parameter locators put on ``garbage_queue`` while the body runs
    (ref: photon/utils.py:73-144 frees Ray ObjectRefs this way; here the queue carries ``ParamHandle`` s of any side channel —
    shm segment, npz object, S3 key — and freeing is ``release_remote_parameters``). ``None`` on the queue stops the collector."""
    import contextlib
    import queue as _queue
    import threading

    from photon_b200.server.s3_utils import release_remote_parameters

    @contextlib.contextmanager
    def _cm() -> Any:
        stop = threading.Event()

        def loop() -> None:
            while not stop.is_set() or not garbage_queue.empty():
                try:
                    h = garbage_queue.get(timeout=0.05)
                except _queue.Empty:
                    continue
                if h is None:
                    return
                release_remote_parameters(h)

        t = threading.Thread(target=loop, name="photon-param-gc", daemon=True)
        t.start()
        if list_of_threads is not None:
            list_of_threads.append(t)
        try:
            yield
        finally:
            stop.set()
            if join_at_the_end:
                t.join(timeout=timeout)

    return _cm()

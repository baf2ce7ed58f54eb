"""Client-side config surgery: turn the resolved ``llm_config`` into a per-client
training config (streams, save/load paths, batch geometry, logger names).

Behavioural parity with the reference's helpers (ref: photon/clients/
llm_config_functions.py:58-1109), written against our ConfigNode tree:

* streams: client ``cid`` owns entry ``streams[cid % len]``; ``cid=None``
  concatenates every client's streams; ``split_eval`` makes one eval loader per
  client labelled ``client_{i}`` (``:239-529``);
* dataset defaults: ``predownload = 8·batch``, ``num_canonical_nodes = 64``,
  ``shuffle_block_size = max(4e6/64, 2^18)`` (``:532-606``);
* per-client ``save_folder = {save_folder}/client_{cid}`` and the mid-round
  resume / skip decision table (``:609-764``);
* ``global_train_batch_size`` rounded down to a multiple of the GPU count
  (``:865-900``); ``num_workers: auto`` = cores / GPUs capped at 32 (``:903-968``);
* unigram frequency dicts merged over the client's streams, cached under /tmp
  (``:971-1109``).
"""
from __future__ import annotations

import copy
import dataclasses
import json
import os
import re
from pathlib import Path
from typing import Any

from photon_b200.config.composer import ConfigNode
from photon_b200.data.synthetic import SyntheticC4

_STREAM_KEYS = ("remote", "local", "proportion", "repeat", "choose", "download_retry", "download_timeout",
                "validate_hash", "keep_zip", "split")


@dataclasses.dataclass
class StreamDict:
    """One entry of ``train_loader.dataset.streams`` (the keyword arguments of a mosaicml-streaming ``Stream``; ref:
    llm_config_functions.py:42-55). ``data/streaming.py`` reads ``local`` / ``remote`` / ``split`` / ``proportion`` / ``repeat`` /
    ``choose``; the download knobs are accepted for files written for the reference."""

    remote: str | None = None
    local: str | None = None
    split: str | None = None
    proportion: float | None = None
    repeat: float | None = None
    choose: int | None = None
    download_retry: int | None = None
    download_timeout: float | None = None
    validate_hash: str | None = None
    keep_zip: bool | None = None

    def to_dict(self) -> dict[str, Any]:
        return {k: v for k, v in dataclasses.asdict(self).items() if v is not None}



def get_train_config(llm_config: Any, *, run_uuid: str | None = None) -> ConfigNode:
    """Deep copy of ``llm_config`` as the mutable per-client TrainConfig, with llm-foundry's
    implicit defaults made explicit."""
    t = ConfigNode(copy.deepcopy(dict(llm_config)))
    t.setdefault("run_name", run_uuid or os.environ.get("RUN_NAME", "photon-b200"))
    t.setdefault("eval_subset_num_batches", -1)
    t.setdefault("eval_first", False)
    t.setdefault("save_num_checkpoints_to_keep", -1)
    t.setdefault("save_overwrite", False)
    t.setdefault("load_path", None)
    t.setdefault("device_eval_batch_size", t.get("device_train_microbatch_size", 1))
    if isinstance(t.get("save_folder"), str):
        t["save_folder"] = t["save_folder"].replace("{run_name}", str(t["run_name"]))
    return t


# ----------------------------------------------------------------------------- streams
def preprocess_stream_paths(dataset_cfg: Any) -> tuple[str | None, str | None, str | None]:
    """(root_local, root_remote, split) and strip them from the dataset node."""
    root_local = dataset_cfg.pop("root_local", None)
    root_remote = dataset_cfg.pop("root_remote", None)
    split = dataset_cfg.pop("split", None)
    return root_local, root_remote, split


def _qualify(stream: dict[str, Any], root_local: str | None, root_remote: str | None, split: str | None) -> dict[str, Any]:
    s = {k: stream.get(k) for k in _STREAM_KEYS if k in stream}
    local, remote = s.get("local"), s.get("remote")
    if root_local is not None:
        s["local"] = os.path.join(root_local, local) if local else root_local
    if root_remote is not None:
        s["remote"] = os.path.join(root_remote, remote) if remote else root_remote
    if split is not None and not s.get("split"):
        s["split"] = split
    return s


def concatenate_streams(clients_streams: list[dict[str, Any]]) -> dict[str, Any]:
    """Union of every client's streams; name clashes get a ``c{i}_`` prefix."""
    out: dict[str, Any] = {}
    for i, entry in enumerate(clients_streams):
        for name, st in (entry.get("client_streams") or {}).items():
            key = name if name not in out else f"c{i}_{name}"
            out[key] = st
    return out


def get_actual_stream(streams: list[dict[str, Any]], cid: int | str | None) -> dict[str, Any]:
    if cid is None:
        return concatenate_streams(streams)
    entry = streams[int(cid) % len(streams)]
    return dict(entry.get("client_streams") or {})


def set_stream(dataset_cfg: Any, cid: int | str | None) -> None:
    """Replace ``dataset.streams`` (list over clients) by this client's ``{name: stream}`` map."""
    streams = dataset_cfg.get("streams")
    if not streams:
        return
    root_local, root_remote, split = preprocess_stream_paths(dataset_cfg)
    chosen = get_actual_stream([dict(s) for s in streams], cid)
    dataset_cfg["streams"] = {n: _qualify(dict(s or {}), root_local, root_remote, split) for n, s in chosen.items()}
    dataset_cfg["split"] = None


def get_split_streams(dataset_cfg: Any) -> dict[str, dict[str, Any]]:
    """One dataset config per client, labelled ``client_{i}`` (``fl.split_eval``)."""
    streams = [dict(s) for s in (dataset_cfg.get("streams") or [])]
    out = {}
    for i in range(len(streams)):
        sub = ConfigNode(copy.deepcopy(dict(dataset_cfg)))
        set_stream(sub, i)
        out[f"client_{i}"] = sub
    return out


def client_set_data_config(train_cfg: Any, cid: int | str | None, split_eval: bool = False) -> dict[str, Any]:
    """Apply stream selection to train/eval loaders; returns ``{label: eval_loader_cfg}``."""
    set_stream(train_cfg["train_loader"]["dataset"], cid)
    eval_loader = train_cfg.get("eval_loader")
    evals: dict[str, Any] = {}
    if eval_loader is not None:
        if split_eval:
            for label, ds in get_split_streams(eval_loader["dataset"]).items():
                lc = ConfigNode(copy.deepcopy(dict(eval_loader)))
                lc["dataset"] = ds
                lc["label"] = label
                evals[label] = lc
        else:
            set_stream(eval_loader["dataset"], None)  # eval always sees all streams concatenated
            evals["eval"] = eval_loader
    return evals


def set_dataset_default_params(train_cfg: Any) -> None:
    bs = int(train_cfg.get("global_train_batch_size", 1))
    for key in ("train_loader", "eval_loader"):
        ld = train_cfg.get(key)
        if not ld:
            continue
        ds = ld["dataset"]
        if ds.get("predownload") is None:
            ds["predownload"] = 8 * bs
        if ds.get("num_canonical_nodes") is None:
            ds["num_canonical_nodes"] = 64
        if ds.get("shuffle_block_size") is None:
            ds["shuffle_block_size"] = int(max(4_000_000 // int(ds["num_canonical_nodes"]), 1 << 18))


# ------------------------------------------------------------------- save / load paths
_CKPT_RE = r"ep(\d+)-ba(\d+)-rank\d+\.pt$"


def set_client_save_and_load_path(train_cfg: Any, cid: int | str | None) -> None:
    if train_cfg.get("save_folder") is not None and cid is not None:
        train_cfg["save_folder"] = str(Path(str(train_cfg["save_folder"])) / f"client_{cid}")
    train_cfg["load_path"] = None


def list_client_checkpoints(folder: str | os.PathLike) -> list[tuple[int, int]]:
    p = Path(folder)
    if not p.is_dir():
        return []
    found = {(int(m.group(1)), int(m.group(2))) for f in p.iterdir() if (m := re.search(_CKPT_RE, f.name))}
    return sorted(found, key=lambda eb: eb[1])


def set_client_load_path(train_cfg: Any, cid: int | str | None, n_steps: int) -> tuple[bool, bool]:
    """Mid-round resume / skip decision (ref: llm_config_functions.py:642-764).

    ``n_steps`` = ``server_steps_cumulative + local_steps`` — the batch count this round must
    END at.  Returns ``(skip_iteration, load_path_set)``:

    * no checkpoints → (False, False): train from the broadcast weights;
    * a checkpoint with ``ba == n_steps`` exists → (True, True): the round was already
      completed before a crash; load it, do not train, do not save again;
    * else if the latest checkpoint has ``ba < n_steps`` → (False, True): resume from it;
    * else (only newer checkpoints) → (False, False).
    """
    set_client_save_and_load_path(train_cfg, cid)
    folder = train_cfg.get("save_folder")
    if folder is None:
        return False, False
    pairs = list_client_checkpoints(folder)
    if not pairs:
        return False, False
    exact = next(((e, b) for e, b in pairs if b == n_steps), None)
    if exact is not None:
        train_cfg["load_path"] = str(Path(folder) / f"ep{exact[0]}-ba{exact[1]}-rank{{rank}}.pt")
        train_cfg["save_folder"] = None
        return True, True
    e, b = pairs[-1]
    if b < n_steps:
        train_cfg["load_path"] = str(Path(folder) / f"ep{e}-ba{b}-rank{{rank}}.pt")
        return False, True
    return False, False


# ------------------------------------------------------------------------ logger names
def set_client_loggers(train_cfg: Any, log_name: str) -> None:
    """Per-client wandb run ``{run}{log_name}`` / tensorboard dir (ref: :767-862)."""
    loggers = train_cfg.get("loggers") or {}
    if "wandb" in loggers:
        kw = dict((loggers["wandb"] or {}).get("init_kwargs") or {})
        for k in ("name", "id"):
            if kw.get(k) is not None:
                kw[k] = f"{kw[k]}{log_name}"
        loggers["wandb"]["init_kwargs"] = kw
    if "tensorboard" in loggers:
        loggers["tensorboard"] = dict(loggers["tensorboard"] or {})
    train_cfg["run_name"] = f"{train_cfg.get('run_name', 'run')}{log_name}"


def set_client_wandb_logger(train_cfg: Any, log_name: str) -> None:
    """Suffix the wandb run name / id with the client's log name (ref: llm_config_functions.py:767-815)."""
    loggers = train_cfg.get("loggers") or {}
    if "wandb" in loggers:
        kw = dict((loggers["wandb"] or {}).get("init_kwargs") or {})
        for k in ("name", "id"):
            if kw.get(k) is not None:
                kw[k] = f"{kw[k]}{log_name}"
        loggers["wandb"] = {**dict(loggers["wandb"] or {}), "init_kwargs": kw}


def set_client_tensorboard_logger(train_cfg: Any, log_name: str) -> None:
    """Give the client its own tensorboard directory: the run name carries the log name (ref: :818-862)."""
    loggers = train_cfg.get("loggers") or {}
    if "tensorboard" in loggers:
        loggers["tensorboard"] = dict(loggers["tensorboard"] or {})
    if not str(train_cfg.get("run_name", "run")).endswith(log_name):
        train_cfg["run_name"] = f"{train_cfg.get('run_name', 'run')}{log_name}"


def set_icl_tasks_root_dir(icl_tasks_listconfig: list[dict[str, Any]], root_dir: str) -> None:
    """Re-root every task's ``dataset_uri`` under ``root_dir`` in place (ref: :202-236). ``icl_tasks_config.root_dir`` does the same
    at evaluation time without touching the task table."""
    for task in icl_tasks_listconfig:
        task["dataset_uri"] = f"{root_dir}/{task['dataset_uri']}"


# ---------------------------------------------------------------------- batch geometry
def adapt_train_batch_size_to_num_devices(train_cfg: Any, n_devices: int) -> None:
    n = max(1, int(n_devices))
    gbs = int(train_cfg["global_train_batch_size"])
    if gbs % n:
        new = max(n, gbs // n * n)
        print(f"[config] global_train_batch_size {gbs} -> {new} (multiple of {n} devices)")
        train_cfg["global_train_batch_size"] = new


def set_n_workers_dataloaders(train_cfg: Any, n_devices: int, n_cpu_cores: int | None = None) -> None:
    cores = n_cpu_cores or os.cpu_count() or 1
    auto = max(1, min(32, cores // max(1, n_devices)))
    for key in ("train_loader", "eval_loader"):
        ld = train_cfg.get(key)
        if ld and ld.get("num_workers") in ("auto", None):
            ld["num_workers"] = auto


# ------------------------------------------------------------------- unigram frequencies
def merge_freq_dicts(dicts: list[dict[str, int]]) -> dict[str, int]:
    out: dict[str, int] = {}
    for d in dicts:
        for k, v in d.items():
            out[str(k)] = out.get(str(k), 0) + int(v)
    return out


def get_stream_freq_dict_for_client(train_cfg: Any, cid: int | str | None, split: str = "train",
                                    allow_failures: bool = False, cache_dir: str = "/tmp") -> dict[str, int] | None:
    """Merge ``{stream}/{split}/1_gram.json`` over the client's streams; synthetic streams use
    the generator's exact distribution. Cached as ``{cache_dir}/{cid}_freq_dict.json``."""
    streams = (train_cfg["train_loader"]["dataset"].get("streams") or {})
    vocab = int((train_cfg.get("model") or {}).get("vocab_size", 0) or 0)
    # the reference keys this cache by client id only (``/tmp/{cid}_freq_dict.json``), so a second experiment on the same box
    # silently reuses the first one's table; here the key also covers WHAT was counted
    import hashlib

    sig = hashlib.sha1(json.dumps([sorted((n, str((s or {}).get("local")), str((s or {}).get("split") or split)) for n, s in streams.items()),
                                   vocab]).encode()).hexdigest()[:10]  # noqa: S324
    cache = Path(cache_dir) / f"{cid}_{sig}_freq_dict.json"
    if cache.exists():
        return json.loads(cache.read_text())
    dicts = []
    for name, st in streams.items():
        local = (st or {}).get("local")
        sp = (st or {}).get("split") or split
        path = Path(str(local)) / sp / "1_gram.json" if local else None
        if path is not None and path.exists():
            dicts.append(json.loads(path.read_text()))
        elif local is None or str(local).startswith("synthetic://") or not Path(str(local)).exists():
            from photon_b200.data.synthetic import TOKENIZER_VOCAB

            # the stand-in stream is generated inside the model's vocabulary (see build_text_loader): count the same thing
            p = SyntheticC4(vocab_size=min(vocab, TOKENIZER_VOCAB) if vocab else TOKENIZER_VOCAB).unigram_probabilities()
            dicts.append({str(i): int(round(x * 1e9)) for i, x in enumerate(p) if x > 0})
        elif not allow_failures:
            raise FileNotFoundError(f"1_gram.json missing for stream '{name}' at {path}")
    if not dicts:
        return None
    merged = merge_freq_dicts(dicts)
    try:
        cache.write_text(json.dumps(merged))
    except OSError:
        pass
    return merged

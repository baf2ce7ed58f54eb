"""Client-side round plumbing around the Trainer: payload decoding, optimizer-state
injection, personalised / randomly re-initialised layers, result post-processing,
initial-parameter construction (ref: photon/clients/utils.py:145-1008).

Payloads are FLAT tensors in the model's sorted-name layout (one plane, or three
planes ``[params | exp_avg | exp_avg_sq]`` when ``fl.aggregate_momenta``); lists of
ndarrays (the reference's representation) are accepted and produced at the
boundaries for the shm / file transports.
"""
from __future__ import annotations

import math
import time
from typing import Any, Sequence

import numpy as np
import torch

from photon_b200.clients.configs import FitConfig
from photon_b200.messages import ClientState
from photon_b200.models.mpt import MPTConfig, MPTForCausalLM
from photon_b200.train.trainer import Trainer
from photon_b200.utils.core import load_model_parameters_from_file
from photon_b200.utils.flat import FlatLayout

Payload = torch.Tensor | Sequence[np.ndarray]


def get_client_state_struct(fit_config: FitConfig, cid: int | str) -> ClientState:
    return fit_config.state_of(int(cid))


def load_ignore_keys(fit_config: FitConfig) -> list[str]:
    """Glob paths dropped when loading a client checkpoint (ref: clients/utils.py:229-238)."""
    keys = ["*scheduler*"]
    if fit_config.reset_optimizer:
        keys.append("*optim*")
    if fit_config.reset_dataset_state:
        keys.append("*dataset_state*")
    if fit_config.reset_timestamp:
        keys.append("*timestamp*")
    return keys


def set_initial_config_from_fit_config(fit_config: FitConfig, llm_config: Any, cid: int | str | None
                                       ) -> tuple[bool, bool, int | None]:
    """Per-round, per-client config surgery before the trainer is (re)configured
    (ref: clients/utils.py:177-254): choose the save folder / resume checkpoint (or decide the round was
    already done), set the load/save ignore globs, apply ``resize_vocab``.

    Returns ``(skip_iteration, checkpoint_exists, server_steps_cumulative)``."""
    from photon_b200.clients import llm_config_functions as lcf
    from photon_b200.train.timestamp import Time

    local_steps = Time.parse(llm_config.get("local_steps", "1ba")).to_batches()
    server_steps = fit_config.server_steps_cumulative
    skip_iteration = checkpoint_exists = False
    if not fit_config.reset_checkpoint:
        if server_steps is None:
            raise ValueError("Server steps cumulative is None and we want to reset a checkpoint.")
        skip_iteration, checkpoint_exists = lcf.set_client_load_path(llm_config, cid, int(server_steps) + local_steps)
    else:
        lcf.set_client_save_and_load_path(llm_config, cid)
    llm_config["load_ignore_keys"] = load_ignore_keys(fit_config)
    if fit_config.reset_optimizer:
        llm_config["save_ignore_keys"] = ["*optim*"]
    model = llm_config.get("model")
    if model is not None and "vocab_size" in model and fit_config.resize_vocab is not None:
        model["vocab_size"] = int(fit_config.resize_vocab)
    return skip_iteration, checkpoint_exists, server_steps


# ------------------------------------------------------------------------ payload codecs
def payload_to_planes(payload: Payload, layout: FlatLayout, device: torch.device, n_planes: int) -> list[torch.Tensor]:
    """→ ``n_planes`` flat fp32 tensors on ``device`` in ``layout`` order."""
    total = layout.total
    if torch.is_tensor(payload):
        flat = payload.reshape(-1)
        if flat.numel() != n_planes * total:
            raise ValueError(f"flat payload has {flat.numel()} elements, expected {n_planes}×{total}")
        return [flat[i * total:(i + 1) * total].to(device, torch.float32) for i in range(n_planes)]
    arrays = list(payload)
    n = len(layout.names)
    if len(arrays) != n_planes * n:
        raise ValueError(f"payload has {len(arrays)} arrays, expected {n_planes}×{n}")
    planes = []
    for i in range(n_planes):
        t = torch.zeros(total, dtype=torch.float32, device=device)
        layout.from_ndarrays(t, arrays[i * n:(i + 1) * n])
        planes.append(t)
    return planes


def planes_to_ndarrays(planes: Sequence[torch.Tensor], layout: FlatLayout) -> list[np.ndarray]:
    out: list[np.ndarray] = []
    for p in planes:
        out.extend(layout.to_ndarrays(p))
    return out


# ------------------------------------------------------------------- optimizer state sync
def set_optimizer_state(trainer: Trainer, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
                        personalized: Sequence[str] = ()) -> dict[str, float]:
    """Overwrite the local Adam/ADOPT moments + step with the server's aggregate
    (``fl.aggregate_momenta``; ref: clients/utils.py:257-402). Personalised layers keep theirs."""
    opt, lay = trainer.state.optimizer, trainer.state.flat.layout
    if personalized:
        cur_m, cur_v = opt.full_moments()
        for i, name in enumerate(lay.names):
            if any(p in name for p in personalized):
                lay.view(exp_avg, i).copy_(lay.view(cur_m, i))
                lay.view(exp_avg_sq, i).copy_(lay.view(cur_v, i))
    opt.set_full_moments(exp_avg, exp_avg_sq.clamp_(min=0.0))   # sharded state keeps only this rank's slice
    opt.step_count = int(step)
    name = opt.name
    return {f"client/local_{name}/l2_norm_exp_avg": float(exp_avg.norm()),
            f"client/local_{name}/l2_norm_exp_avg_sq": float(exp_avg_sq.norm()),
            f"client/local_{name}/step": float(step)}


def personalize_layers(incoming: torch.Tensor, local: torch.Tensor, layout: FlatLayout, personalized: Sequence[str]) -> int:
    """Keep the client's LOCAL values for every tensor whose name contains one of
    ``personalized`` (ref: clients/utils.py:950-1008). Returns the number of tensors kept."""
    kept = 0
    for i, name in enumerate(layout.names):
        if any(p in name for p in personalized):
            layout.view(incoming, i).copy_(layout.view(local, i))
            kept += 1
    return kept


def randomize_layers(incoming: torch.Tensor, layout: FlatLayout, model_cfg: MPTConfig, random_layers: Sequence[str], *,
                     server_round: int, cid: int, truly_random: bool, base_seed: int = 17) -> int:
    """Replace the listed tensors by a freshly initialised model's values
    (ref: clients/utils.py:871-947). ``truly_random`` seeds by (round, cid) so every
    re-init differs; otherwise the same ``base_seed`` init is re-applied."""
    seed = (base_seed + 7919 * int(server_round) + 104729 * int(cid)) if truly_random else base_seed
    fresh = MPTForCausalLM(model_cfg, device="cpu", seed=seed)
    named = dict(fresh.named_parameters())
    done = 0
    for i, name in enumerate(layout.names):
        if any(r in name for r in random_layers):
            layout.view(incoming, i).copy_(named[name].detach().to(incoming.device))
            done += 1
    return done


def manipulate_pre_training_params(trainer: Trainer, payload: Payload, fit_config: FitConfig, cid: int,
                                   client_state: ClientState) -> tuple[torch.Tensor, dict[str, float]]:
    """Decode the round payload and apply momenta / personalisation / random re-init
    (ref: clients/utils.py:405-511). Returns the flat params to install + metrics."""
    st = trainer.state
    lay, dev = st.flat.layout, st.flat.params.device
    n_planes = 3 if fit_config.aggregate_momenta else 1
    planes = payload_to_planes(payload, lay, dev, n_planes)
    params = planes[0].clone() if torch.is_tensor(payload) and planes[0].data_ptr() == payload.data_ptr() else planes[0]
    metrics: dict[str, float] = {}
    pers = list(fit_config.personalized_layers or [])
    # key filter (ref: photon/utils.py:640-670): only tensors whose name contains ``fl.set_trainer_key_to_filter`` are taken
    # from the server. The flat payload always carries every tensor, so the filtered-out ones simply keep the client's local
    # values — the same thing a reference client sees when those tensors never travel. ("transformer" matches all of MPT.)
    if fit_config.set_trainer_params_filter_keys and fit_config.set_trainer_key_to_filter:
        pers += [n for n in lay.names if fit_config.set_trainer_key_to_filter not in n]
    if fit_config.aggregate_momenta:
        metrics.update(set_optimizer_state(trainer, planes[1], planes[2], client_state.local_steps_cumulative, pers))
    if pers and client_state.local_steps_cumulative > 0:
        personalize_layers(params, st.flat.params, lay, pers)
    rl, freq = list(fit_config.random_layers or []), int(fit_config.random_init_freq or 0)
    if rl and freq > 0 and client_state.local_steps_cumulative % freq == 0:
        randomize_layers(params, lay, trainer.model_cfg, rl, server_round=fit_config.server_round, cid=cid,
                         truly_random=fit_config.truly_random_init, base_seed=trainer.seed)
    return params, metrics


def manipulate_pre_training_ndarrays(parameters: Sequence[np.ndarray], trainer: Trainer, fit_config: FitConfig,
                                     client_state: ClientState, cid: int = 0) -> list[np.ndarray]:
    """The reference's ndarray-in / ndarray-out form of ``manipulate_pre_training_params``
    (ref: clients/utils.py:405-511): momenta are split off and installed into the optimizer, personalised /
    randomly re-initialised layers are patched, and only the model's arrays come back."""
    params, _ = manipulate_pre_training_params(trainer, list(parameters), fit_config, cid, client_state)
    return trainer.state.flat.layout.to_ndarrays(params)


# ------------------------------------------------------------------------- post-processing
def post_process_client_result(trainer: Trainer, initial: torch.Tensor, fit_config: FitConfig, cid: int,
                               client_state: ClientState, steps_done: int, as_ndarrays: bool = False
                               ) -> tuple[Payload, int, dict[str, Any]]:
    """Pack the result payload, n_samples and the client metrics (ref: clients/utils.py:514-652)."""
    t0 = time.time_ns()
    st = trainer.state
    lay = st.flat.layout
    n_samples = max(1, int(steps_done) * int(fit_config.batch_size))  # ref: clients/utils.py:583
    metrics: dict[str, Any] = dict(st.train_metric_values)
    # per-tensor pseudo-gradient norms with ONE host sync and no parameter-sized fp64 temporaries (the reference does one
    # NumPy reduction per tensor on the CPU, ref: clients/utils.py:599-619): a multi-tensor norm over views of the delta
    delta = initial - st.flat.params
    views = [lay.view(delta, i) for i in range(len(lay.names))]
    per = torch.stack(torch._foreach_norm(views)).double()
    vals = torch.cat([per, (per * per).sum().sqrt()[None]]).tolist()
    del delta, views
    metrics["client/l2_norm_pseudo_gradient"] = vals[-1]
    for i in range(len(lay.names)):
        metrics[f"client/layer/{i}/l2_norm_of_pseudo_gradient"] = vals[i]
    planes = [st.flat.params]
    if fit_config.aggregate_momenta:
        planes += list(st.optimizer.full_moments())
    new_state = ClientState(local_steps_cumulative=client_state.local_steps_cumulative + int(steps_done),
                            local_timestamp={k: v for k, v in st.timestamp.state_dict().items() if isinstance(v, (int, float))},
                            steps_done=int(steps_done))
    metrics["client_state_acc"] = str({int(cid): new_state.to_literal()})
    payload: Payload = planes_to_ndarrays(planes, lay) if as_ndarrays else (
        torch.cat([p.reshape(-1) for p in planes]) if len(planes) > 1 else planes[0])
    metrics["client/fit_metrics_collection_time"] = (time.time_ns() - t0) / 1e9
    return payload, n_samples, metrics


# --------------------------------------------------------------------- initial parameters
def get_raw_model_parameters(cfg: Any, *, with_momenta: bool = False, seed: int | None = None) -> tuple[list[np.ndarray], FlatLayout]:
    """Build the model on CPU and return its trainable tensors in sorted-name order, optionally
    followed by two zero planes (momenta) (ref: clients/utils.py:739-868)."""
    model_node = dict(cfg["llm_config"]["model"])
    fl = cfg.get("fl") or {}
    if fl.get("resize_vocab"):
        model_node["vocab_size"] = int(fl["resize_vocab"])
    mcfg = MPTConfig.from_model_cfg(model_node)
    model = MPTForCausalLM(mcfg, device="cpu", seed=int(cfg["llm_config"].get("seed", 17)) if seed is None else seed)
    from photon_b200.train.backend import apply_freeze

    apply_freeze(model, fl.get("frozen_layers"), fl.get("unfrozen_layers"))
    named = sorted(((n, p) for n, p in model.named_parameters() if p.requires_grad), key=lambda kv: kv[0])
    layout = FlatLayout.build((n, p.shape) for n, p in named)
    arrays = [p.detach().numpy().copy() for _, p in named]
    if with_momenta:
        arrays += [np.zeros_like(a) for a in arrays] + [np.zeros_like(a) for a in arrays]
    return arrays, layout


def get_initial_parameters(cfg: Any) -> tuple[list[np.ndarray], FlatLayout]:
    """Fresh init or ``pretrained_model_path`` (npz/bin) with a parameter-count check
    (ref: clients/utils.py:676-736)."""
    arrays, layout = get_raw_model_parameters(cfg, with_momenta=False)
    path = cfg.get("pretrained_model_path")
    if path:
        loaded = load_model_parameters_from_file(path)
        n = len(layout.names)
        if len(loaded) < n:
            raise AssertionError(f"pretrained model has {len(loaded)} tensors, model needs {n}")
        for a, b, name in zip(loaded[:n], arrays, layout.names):
            if a.shape != b.shape:
                raise AssertionError(f"pretrained tensor for {name}: {a.shape} vs {b.shape}")
        arrays = [np.asarray(a, dtype=np.float32) for a in loaded[:n]]
    return arrays, layout


def pseudo_gradient_norm(initial: torch.Tensor, final: torch.Tensor) -> float:
    return float(math.sqrt(float(((initial - final).double() ** 2).sum())))


def streaming_shms_clean_up(prefixes: Sequence[str] = ("photon_", "pb200_", "nm-"), min_age_s: float = 0.0) -> int:
    """Unlink stale POSIX shared-memory segments left by crashed node managers / workers / loaders and collect garbage
    (ref: clients/utils.py:655-673, which leans on streaming's stale-shm sweep). ``min_age_s`` spares segments younger
    than that (another federation may be starting on the same box): ``python -m photon_b200.clients.utils 3600`` removes
    what has not been touched for an hour."""
    import gc
    import os
    import time as _time

    removed = 0
    try:
        names = os.listdir("/dev/shm")
    except OSError:
        names = []
    from photon_b200.shm.utils import unlink_quietly

    now = _time.time()
    for n in names:
        if not n.startswith(tuple(prefixes)):
            continue
        try:
            if min_age_s and now - os.stat(os.path.join("/dev/shm", n)).st_mtime < min_age_s:
                continue
        except OSError:
            continue
        unlink_quietly(n)
        removed += 1
    gc.collect()
    return removed


if __name__ == "__main__":
    import sys

    print(f"[shm] removed {streaming_shms_clean_up(min_age_s=float(sys.argv[1]) if len(sys.argv) > 1 else 600.0)} stale segment(s)")

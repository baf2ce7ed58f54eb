"""``llm_fit`` / ``llm_eval`` — one federated client's round on a (re-used) Trainer.

Same contract and step order as the reference (ref: photon/clients/
llm_client_functions.py:53-228 fit, :231-353 eval), with tensors staying on the
device: the incoming global model is a flat device tensor (or the NVLink arena's
global plane) and the result is the Trainer's flat parameter plane — no
``.cpu().numpy()`` per tensor, no ndarray lists unless a host transport asks.

Timings reported as metrics keep the reference's names (``client/fit_init_time``,
``client/fit_set_parameters_time``, ``client/fit_time`` …).
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Any

import torch

from photon_b200.clients.configs import EvaluateConfig, FitConfig
from photon_b200.clients.trainer_utils import get_trainer_object, load_trainer_checkpoint, reconfigure_trainer
from photon_b200.clients.utils import (Payload, load_ignore_keys, manipulate_pre_training_params, payload_to_planes,
                                       post_process_client_result, set_initial_config_from_fit_config)
from photon_b200.train.timestamp import Time
from photon_b200.train.trainer import Trainer


def _now() -> float:
    return time.time_ns() / 1e9


def llm_fit(trainer: Trainer | None, payload: Payload, fit_config: FitConfig | dict[str, Any], cfg: Any, cid: int, *,
            as_ndarrays: bool = False, trainer_kwargs: dict[str, Any] | None = None,
            shadow_payload: torch.Tensor | None = None) -> tuple[Payload, int, dict[str, Any], Trainer]:
    """Run client ``cid``'s local training. Returns ``(payload, n_samples, metrics, trainer)``."""
    t_start = _now()
    fc = fit_config if isinstance(fit_config, FitConfig) else FitConfig.from_wire(fit_config)
    state = fc.state_of(cid)
    llm = cfg["llm_config"]
    local_steps = Time.parse(llm.get("local_steps", "1ba")).to_batches()  # NB: llm_config.local_steps, not fl.n_local_steps
    server_steps = int(fc.server_steps_cumulative or 0)
    metrics: dict[str, Any] = {}

    # ---- trainer: build once, afterwards only swap the per-client mutables
    if trainer is None:
        trainer, train_cfg = get_trainer_object(cfg, cid, log_name=f"_client_{cid}", split_eval=fc.split_eval,
                                                use_unigram_metrics=fc.use_unigram_metrics,
                                                allow_unigram_metrics_failures=fc.allow_unigram_metrics_failures,
                                                frozen_layers=fc.frozen_layers, unfrozen_layers=fc.unfrozen_layers,
                                                resize_vocab=fc.resize_vocab, **(trainer_kwargs or {}))
    else:
        train_cfg = reconfigure_trainer(trainer, cfg, cid, log_name=f"_client_{cid}", split_eval=fc.split_eval,
                                        use_unigram_metrics=fc.use_unigram_metrics,
                                        allow_unigram_metrics_failures=fc.allow_unigram_metrics_failures)
    # ---- per-client checkpoint policy: resume mid-round or skip an already finished round. With an object store configured the
    # client's newest checkpoint is fetched first: it may have been written on another machine (cross-host node fleet)
    from photon_b200.checkpoint.store import client_checkpoint_mirror

    mirror = client_checkpoint_mirror(cfg) if (train_cfg.get("save_folder") and not fc.reset_checkpoint) else None
    mirror_folder = str(Path(str(train_cfg["save_folder"])) / f"client_{cid}") if mirror is not None else None

    def _ranks_together() -> None:      # the ranks of one client share the folder: nobody lists / uploads it half way
        if trainer.world_size > 1 and torch.distributed.is_initialized():
            torch.distributed.barrier(group=trainer.process_group)

    if mirror is not None:
        if trainer.rank == 0:
            fetched = mirror.pull(cid, mirror_folder)
            if fetched:
                metrics["client/checkpoints_fetched"] = len(fetched)
        _ranks_together()
    skip_iteration, load_set, _ = set_initial_config_from_fit_config(fc, train_cfg, cid)
    trainer.save_folder = train_cfg.get("save_folder")
    trainer.save_ignore_keys = list(train_cfg.get("save_ignore_keys") or [])   # reset_optimizer → client checkpoints carry no optimizer state
    metrics["client/fit_init_time"] = _now() - t_start

    # ---- client checkpoint FIRST (optimizer moments, data position, and the model it holds), THEN the round's parameters on
    # top — the reference's order (ref: llm_client_functions.py:126-168). Only a client that already finished this round
    # (skip_iteration) keeps the checkpoint's model: it IS the round's result. Loading after the install would overwrite the
    # freshly broadcast global model with the client's stale weights and silently bypass the aggregation.
    t0 = _now()
    st = trainer.state
    if load_set and train_cfg.get("load_path"):
        load_trainer_checkpoint(trainer, str(train_cfg["load_path"]), load_ignore_keys(fc))
    if fc.reset_optimizer and not fc.aggregate_momenta:
        st.optimizer.reset_state()
    params, m = manipulate_pre_training_params(trainer, payload, fc, cid, state)
    metrics.update(m)
    if not skip_iteration:
        st.flat.params.copy_(params)
        if shadow_payload is not None and getattr(st.backend, "bf16_params", None) is not None:
            st.backend.bf16_params.copy_(shadow_payload)  # bf16 cast already produced by the round broadcast kernel
        else:
            st.backend.params_updated()
    initial = params if params.data_ptr() != st.flat.params.data_ptr() else params.clone()
    # LR schedule continuity: the local clock starts at the federation's cumulative step count (a checkpoint written in the
    # middle of the round restarts the round from the global model, like the reference)
    if skip_iteration:
        pass
    elif not fc.reset_timestamp:
        st.timestamp.batch = server_steps
    else:
        st.timestamp.reset()
    metrics["client/fit_set_parameters_time"] = _now() - t0

    if llm.get("eval_first", False) and trainer.eval_loaders:
        t0 = _now()
        pre = trainer.eval()
        metrics.update({f"PrePersonalization{k}": v for k, v in pre.items()})
        metrics["client/fit_pre_eval_time"] = _now() - t0

    # ---- local training
    t0 = _now()
    steps_done = 0
    if not skip_iteration:
        target = server_steps + local_steps
        remaining = max(0, target - st.timestamp.batch)
        if remaining:
            trainer.fit(duration=f"{remaining}ba")
        steps_done = local_steps       # (no device synchronise here: the last batch of fit() read its loss back, which already waited)
    metrics["client/fit_time"] = _now() - t0
    if mirror is not None:
        _ranks_together()
        if trainer.rank == 0:
            metrics["client/checkpoints_uploaded"] = len(mirror.push(cid, mirror_folder))

    t0 = _now()
    out, n_samples, pm = post_process_client_result(trainer, initial, fc, cid, state, steps_done, as_ndarrays=as_ndarrays)
    metrics.update(pm)
    metrics["client/fit_get_parameters_time"] = _now() - t0
    return out, n_samples, metrics, trainer


def llm_eval(trainer: Trainer | None, payload: Payload, eval_config: EvaluateConfig | dict[str, Any], cfg: Any,
             cid: int | None = None, *, trainer_kwargs: dict[str, Any] | None = None) -> tuple[float, int, dict[str, Any], Trainer]:
    """Federated evaluation: all streams concatenated (``cid=None``), parameters checked before/after
    install, ``Val*`` metrics; loss is the real CE here (the reference returns a dummy 0.0, ref: :353)."""
    t_start = _now()
    ec = eval_config if isinstance(eval_config, EvaluateConfig) else EvaluateConfig.from_wire(eval_config)
    metrics: dict[str, Any] = {}
    if trainer is None:
        trainer, _ = get_trainer_object(cfg, None, log_name="_eval", split_eval=ec.split_eval,
                                        use_unigram_metrics=ec.use_unigram_metrics,
                                        allow_unigram_metrics_failures=ec.allow_unigram_metrics_failures,
                                        frozen_layers=ec.frozen_layers, unfrozen_layers=ec.unfrozen_layers,
                                        resize_vocab=ec.resize_vocab, **(trainer_kwargs or {}))
    else:
        reconfigure_trainer(trainer, cfg, None, log_name="_eval", split_eval=ec.split_eval, use_unigram_metrics=ec.use_unigram_metrics,
                            allow_unigram_metrics_failures=ec.allow_unigram_metrics_failures)
    metrics["client/eval_init_time"] = _now() - t_start
    t0 = _now()
    st = trainer.state
    before = st.flat.params.clone()
    lay = st.flat.layout
    if torch.is_tensor(payload):   # with fl.aggregate_momenta the broadcast carries [params | exp_avg | exp_avg_sq]:
        payload = payload.reshape(-1)[: lay.total]          # evaluation only needs the model plane
    else:
        payload = list(payload)[: len(lay.names)]
    (params,) = payload_to_planes(payload, lay, st.flat.params.device, 1)
    st.flat.params.copy_(params)
    st.backend.params_updated()
    if not torch.equal(st.flat.params, params):
        raise AssertionError("evaluation parameters were not installed correctly")
    metrics["client/eval_params_changed"] = float(not torch.equal(before, st.flat.params))
    metrics["client/eval_set_parameters_time"] = _now() - t0
    t0 = _now()
    vals = trainer.eval()
    if trainer.device.type == "cuda":
        torch.cuda.synchronize(trainer.device)
    metrics["client/eval_time"] = _now() - t0
    t0 = _now()
    out = {}
    loss = 0.0
    for k, v in vals.items():
        label, _, name = k.rpartition("/")
        out[f"{label + '/' if label and label != 'eval' else ''}Val{name}"] = v
        if name == "LanguageCrossEntropy":
            loss = float(v)
    metrics.update(out)
    n_samples = int(st.eval_timestamp.sample)
    metrics["client/eval_metrics_collection_time"] = _now() - t0
    # the trainer is persistent here (the reference closes and rebuilds it around every evaluation,
    # ref: llm_client_functions.py:341-351): the key is kept for dashboards that expect it
    metrics["client/eval_trainer_closing_time"] = 0.0
    return loss, max(1, n_samples), metrics, trainer

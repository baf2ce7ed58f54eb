"""Trainer callbacks and loggers.

The reference enables these through llm-foundry's registry
(ref: photon/conf/llm_config/mpt-125m.yaml:98-115: ``speed_monitor``,
``lr_monitor``, ``memory_monitor``, ``runtime_estimator``,
``activation_monitor_full_model``, ``optimizer_monitor``; loggers ``wandb`` /
``tensorboard`` + console).  Metric names follow Composer's so dashboards
carry over (``throughput/tokens_per_sec``, ``throughput/device/mfu``,
``lr-<Opt>/group0``, ``memory/peak_allocated_mem``, ``time/remaining_estimate``,
``l2_norm/grad/global`` …).
"""
from __future__ import annotations

import json
import time
from collections import deque
from pathlib import Path
from typing import TYPE_CHECKING, Any

import torch

from photon_b200.train.timestamp import Time

if TYPE_CHECKING:  # pragma: no cover
    from photon_b200.train.trainer import Trainer


# ----------------------------------------------------------------------------- loggers
class Logger:
    def log_metrics(self, metrics: dict[str, float], step: int) -> None:  # noqa: D401
        raise NotImplementedError

    def log_hparams(self, hp: dict[str, Any]) -> None:
        pass

    def close(self) -> None:
        pass


class InMemoryLogger(Logger):
    def __init__(self) -> None:
        self.data: dict[str, list[tuple[int, float]]] = {}
        self.hparams: dict[str, Any] = {}

    def log_metrics(self, metrics: dict[str, float], step: int) -> None:
        for k, v in metrics.items():
            self.data.setdefault(k, []).append((step, float(v)))

    def log_hparams(self, hp: dict[str, Any]) -> None:
        self.hparams.update(hp)

    def latest(self, key: str, default: float | None = None) -> float | None:
        return self.data[key][-1][1] if key in self.data else default


class ConsoleLogger(Logger):
    def __init__(self, interval: int = 1, prefix: str = "") -> None:
        self.interval, self.prefix = max(1, int(interval)), prefix

    def log_metrics(self, metrics: dict[str, float], step: int) -> None:
        if step % self.interval:
            return
        keep = {k: v for k, v in metrics.items() if "/layer/" not in k}
        body = " ".join(f"{k}={v:.5g}" for k, v in sorted(keep.items()))
        print(f"{self.prefix}[ba={step}] {body}", flush=True)


class ProgressBarLogger(Logger):
    """``progress_bar: true`` — one tqdm bar per run showing the batch counter and the running loss (Composer's
    ``ProgressBarLogger``); needs nothing but a terminal."""

    def __init__(self, total: int | None = None, desc: str = "train") -> None:
        from tqdm import tqdm

        self.bar = tqdm(total=total, desc=desc, unit="ba", dynamic_ncols=True, leave=True)
        self._last = 0

    def log_metrics(self, metrics: dict[str, float], step: int) -> None:
        if step > self._last:
            self.bar.update(step - self._last)
            self._last = step
        if "loss/train/total" in metrics:
            self.bar.set_postfix(loss=f"{float(metrics['loss/train/total']):.4f}", refresh=False)

    def close(self) -> None:
        self.bar.close()


class JSONLLogger(Logger):
    """File sink (one JSON object per call). Stands in for the TensorBoard event file
    when ``tensorboard`` is not importable; path mirrors ``<save>/tensorboard/<run>``."""

    def __init__(self, path: str | Path, flush_interval: int = 10) -> None:
        self.path = Path(path)
        self.path.parent.mkdir(parents=True, exist_ok=True)
        self._f = open(self.path, "a", encoding="utf-8")
        self._n, self.flush_interval = 0, max(1, int(flush_interval))

    def log_metrics(self, metrics: dict[str, float], step: int) -> None:
        self._f.write(json.dumps({"step": step, **{k: float(v) for k, v in metrics.items()}}) + "\n")
        self._n += 1
        if self._n % self.flush_interval == 0:
            self._f.flush()

    def close(self) -> None:
        self._f.close()


class TensorBoardLogger(Logger):
    """Scalars as TensorBoard event files: through ``torch.utils.tensorboard`` when the tensorboard package is there, otherwise through
    the own writer of the same on-disk format (``utils/tbevents.py``) — either way ``tensorboard --logdir`` reads the run."""

    def __init__(self, log_dir: str | Path, flush_interval: int = 10) -> None:
        try:
            from torch.utils.tensorboard import SummaryWriter

            self._w: Any = SummaryWriter(str(log_dir), flush_secs=max(1, int(flush_interval)))
            self._native: Any = None
        except Exception:  # noqa: BLE001 - tensorboard missing in this image
            from photon_b200.utils.tbevents import EventFileWriter

            self._w = None
            self._native = EventFileWriter(log_dir, flush_secs=max(1, int(flush_interval)))

    def log_metrics(self, metrics: dict[str, float], step: int) -> None:
        if self._w is not None:
            for k, v in metrics.items():
                self._w.add_scalar(k, float(v), step)
        else:
            self._native.add_scalars({k: float(v) for k, v in metrics.items()}, step)

    def close(self) -> None:
        (self._w or self._native).close()


class WandBLogger(Logger):
    """wandb sink; silently degrades to offline/no-op when wandb is not importable
    (there is no network here). ``init_kwargs`` = ``wandb.setup`` of the config."""

    def __init__(self, init_kwargs: dict[str, Any] | None = None) -> None:
        self._run = None
        try:
            import wandb  # type: ignore[import-not-found]

            kw = dict(init_kwargs or {})
            kw.setdefault("mode", "offline")
            self._run = wandb.init(**kw)
        except Exception:  # noqa: BLE001
            self._run = None

    def log_metrics(self, metrics: dict[str, float], step: int) -> None:
        if self._run is not None:
            self._run.log(metrics, step=step)

    def close(self) -> None:
        if self._run is not None:
            self._run.finish()


def build_loggers(cfg: dict[str, Any] | None, save_root: str | Path, run_name: str, console_interval: int = 1,
                  log_to_console: bool = True, rank: int = 0, progress_bar: bool = False, total_batches: int | None = None) -> list[Logger]:
    out: list[Logger] = [InMemoryLogger()]
    if rank != 0:
        return out
    if log_to_console:
        out.append(ConsoleLogger(console_interval, prefix=f"{run_name} "))
    if progress_bar:
        out.append(ProgressBarLogger(total_batches, desc=run_name))
    for name, sub in (cfg or {}).items():
        sub = dict(sub or {})
        if name == "tensorboard":
            out.append(TensorBoardLogger(Path(save_root) / "tensorboard" / run_name, sub.get("flush_interval", 10)))
        elif name == "wandb":
            out.append(WandBLogger(sub.get("init_kwargs")))
        elif name in ("jsonl", "file"):
            out.append(JSONLLogger(Path(save_root) / f"{run_name}.metrics.jsonl"))
    return out


# --------------------------------------------------------------------------- callbacks
class Callback:
    def fit_start(self, tr: "Trainer") -> None: ...
    def batch_start(self, tr: "Trainer") -> None: ...
    def after_backward(self, tr: "Trainer") -> None: ...
    def batch_end(self, tr: "Trainer") -> None: ...
    def eval_end(self, tr: "Trainer") -> None: ...
    def fit_end(self, tr: "Trainer") -> None: ...
    def state_dict(self) -> dict[str, Any]:
        return {}

    def load_state_dict(self, sd: dict[str, Any]) -> None: ...


class SpeedMonitor(Callback):
    """Rolling-window throughput; wall time is taken after the trainer's per-batch device
    sync point so tokens/s reflects completed work. MFU uses ``6·N + 12·L·d·S`` FLOPs."""

    def __init__(self, window_size: int = 20, gpu_flops_available: float | None = None) -> None:
        self.window = deque(maxlen=int(window_size) + 1)
        self.gpu_flops = gpu_flops_available
        self.total_train_s = 0.0
        self._t0: float | None = None

    def fit_start(self, tr: "Trainer") -> None:
        self._t0 = time.perf_counter()
        self.window.clear()
        self.window.append((self._t0, tr.state.timestamp.sample, tr.state.timestamp.token, tr.state.timestamp.batch))

    def batch_end(self, tr: "Trainer") -> None:
        now = time.perf_counter()
        ts = tr.state.timestamp
        self.window.append((now, ts.sample, ts.token, ts.batch))
        if self._t0 is not None:
            self.total_train_s = now - self._t0
        if len(self.window) < 2:
            return
        t0, s0, k0, b0 = self.window[0]
        dt = max(now - t0, 1e-9)
        w = max(1, tr.world_size)
        sps, tps, bps = (ts.sample - s0) / dt, (ts.token - k0) / dt, (ts.batch - b0) / dt
        m = {"throughput/batches_per_sec": bps, "throughput/samples_per_sec": sps * w,
             "throughput/tokens_per_sec": tps * w, "throughput/device/samples_per_sec": sps,
             "throughput/device/tokens_per_sec": tps, "time/train": self.total_train_s / 3600.0}
        fpt = tr.flops_per_token()
        if fpt:
            m["throughput/device/flops_per_sec"] = tps * fpt
            m["throughput/flops_per_sec"] = tps * fpt * w
            peak = self.gpu_flops or tr.peak_flops()
            if peak:
                m["throughput/device/mfu"] = tps * fpt / peak
        tr.log(m)


class LRMonitor(Callback):
    def batch_end(self, tr: "Trainer") -> None:
        tr.log({f"lr-{type(tr.state.optimizer).__name__}/group0": tr.state.optimizer.lr})


class MemoryMonitor(Callback):
    def batch_end(self, tr: "Trainer") -> None:
        if tr.device.type != "cuda":
            return
        st = torch.cuda.memory_stats(tr.device)
        gb = 1e9
        tr.log({"memory/current_allocated_mem": st.get("allocated_bytes.all.current", 0) / gb,
                "memory/peak_allocated_mem": st.get("allocated_bytes.all.peak", 0) / gb,
                "memory/current_reserved_mem": st.get("reserved_bytes.all.current", 0) / gb,
                "memory/peak_reserved_mem": st.get("reserved_bytes.all.peak", 0) / gb,
                "memory/alloc_retries": st.get("num_alloc_retries", 0)})


class RuntimeEstimator(Callback):
    def __init__(self, skip_batches: int = 1) -> None:
        self.skip, self._t0, self._b0 = skip_batches, None, 0

    def batch_end(self, tr: "Trainer") -> None:
        ts = tr.state.timestamp
        if self._t0 is None:
            if ts.batch - tr.fit_start_batch >= self.skip:
                self._t0, self._b0 = time.perf_counter(), ts.batch
            return
        done = ts.batch - self._b0
        target = tr.fit_end_batch
        if done > 0 and target is not None:
            rate = (time.perf_counter() - self._t0) / done
            tr.log({"time/remaining_estimate": max(0, target - ts.batch) * rate / 3600.0})


class OptimizerMonitor(Callback):
    """Global (and optionally per-tensor) L2 norms of grads / params / moments."""

    def __init__(self, interval: str | int = "10ba", only_global: bool = True) -> None:
        self.interval = max(1, Time.parse(interval).to_batches())
        self.only_global = only_global

    def after_backward(self, tr: "Trainer") -> None:
        if tr.state.timestamp.batch % self.interval:
            return
        flat, opt = tr.state.flat, tr.state.optimizer
        if getattr(flat, "is_sharded", False):      # every plane is this rank's shard: the norms span all ranks
            gn = flat.comm.grad_norm
            tr.log({"l2_norm/grad/global": float(gn(flat.grads)), "l2_norm/param/global": float(gn(flat.params)),
                    "l2_norm/moment/global": float(gn(opt.exp_avg)), "l2_norm/second_moment_sqrt/global": float(gn(opt.exp_avg_sq.sqrt()))})
            return
        m = {"l2_norm/grad/global": float(flat.grads.norm()), "l2_norm/param/global": float(flat.params.norm()),
             "l2_norm/moment/global": float(opt.exp_avg.norm()),
             "l2_norm/second_moment_sqrt/global": float(opt.exp_avg_sq.sqrt().norm())}
        if not self.only_global:
            for i, n in enumerate(flat.layout.names):
                m[f"l2_norm/grad/{n}"] = float(flat.layout.view(flat.grads, i).norm())
        tr.log(m)


class ActivationMonitorFullModel(Callback):
    """Residual-stream statistics per block (l2 norm, mean, max) every ``interval``."""

    def __init__(self, interval: str | int = "10ba") -> None:
        self.interval = max(1, Time.parse(interval).to_batches())

    def batch_start(self, tr: "Trainer") -> None:
        tr.state.backend.collect_activation_stats = (tr.state.timestamp.batch % self.interval == 0)

    def batch_end(self, tr: "Trainer") -> None:
        stats = getattr(tr.state.backend, "activation_stats", None)
        if stats:
            tr.log({f"activations/{k}": float(v) for k, v in stats.items()})
            tr.state.backend.activation_stats = {}


_CALLBACKS = {"speed_monitor": SpeedMonitor, "lr_monitor": LRMonitor, "memory_monitor": MemoryMonitor,
              "runtime_estimator": RuntimeEstimator, "optimizer_monitor": OptimizerMonitor,
              "activation_monitor_full_model": ActivationMonitorFullModel}


class MemorySnapshot(Callback):
    """Dump ``torch.cuda.memory._snapshot()`` pickles for a window of batches (Composer's
    ``memory_snapshot`` callback; viewable at pytorch.org/memory_viz)."""

    def __init__(self, skip_batches: int = 1, interval: str | int = "3ba", max_entries: int = 100000,
                 folder: str = "memory_snapshots") -> None:
        self.skip, self.interval = int(skip_batches), max(1, Time.parse(interval).to_batches())
        self.max_entries, self.folder = int(max_entries), folder
        self._recording = False

    def batch_start(self, tr: "Trainer") -> None:
        if tr.device.type != "cuda":
            return
        b = tr.state.timestamp.batch - tr.fit_start_batch
        if b == self.skip and not self._recording:
            torch.cuda.memory._record_memory_history(max_entries=self.max_entries)
            self._recording = True

    def batch_end(self, tr: "Trainer") -> None:
        if not self._recording:
            return
        b = tr.state.timestamp.batch - tr.fit_start_batch
        if b >= self.skip + self.interval:
            self.dump(tr, f"rank{tr.rank}_ba{tr.state.timestamp.batch}")
            torch.cuda.memory._record_memory_history(enabled=None)
            self._recording = False

    def dump(self, tr: "Trainer", tag: str) -> Path:
        out = Path(tr.save_folder or ".") / self.folder
        out.mkdir(parents=True, exist_ok=True)
        path = out / f"{tag}_memory_snapshot.pickle"
        torch.cuda.memory._dump_snapshot(str(path))
        return path


class OOMObserver(Callback):
    """Record allocator history from the first batch and dump it when the allocator raises
    an out-of-memory error (Composer's ``oom_observer``)."""

    def __init__(self, max_entries: int = 100000, folder: str = "oom_snapshots") -> None:
        self.max_entries, self.folder = int(max_entries), folder
        self.dumped: list[Path] = []

    def fit_start(self, tr: "Trainer") -> None:
        if tr.device.type != "cuda":
            return
        torch.cuda.memory._record_memory_history(max_entries=self.max_entries)
        out = Path(tr.save_folder or ".") / self.folder

        def observer(device: int, alloc: int, device_alloc: int, device_free: int) -> None:
            out.mkdir(parents=True, exist_ok=True)
            path = out / f"rank{tr.rank}_oom_memory_snapshot.pickle"
            torch.cuda.memory._dump_snapshot(str(path))
            self.dumped.append(path)

        torch._C._cuda_attach_out_of_memory_observer(observer)


class ProfilerCallback(Callback):
    """``llm_config.profiler``: Composer's ``Profiler(schedule=cyclic_schedule(...), trace_handlers=[JSONTraceHandler])``
    surface (ref: clients/trainer_utils.py:1456-1482) on top of ``torch.profiler``: per cycle ``skip_first`` (once),
    ``wait``, ``warmup``, ``active`` batches, ``repeat`` cycles (0 = forever); every finished active window is exported
    as a Chrome-trace JSON ``{folder}/rank{R}.{batch}.pt.trace.json``. CPU + CUDA activities; our kernels show up by
    name (they are ordinary launches), and `PhaseTracer` spans can be correlated through NVTX."""

    def __init__(self, schedule: dict[str, Any] | None = None, json_trace_handler: dict[str, Any] | None = None,
                 torch_prof_record_shapes: bool = False, torch_prof_profile_memory: bool = False,
                 torch_prof_with_stack: bool = False, torch_prof_with_flops: bool = False, **_ignored: Any) -> None:
        sch = dict(schedule or {})
        self.skip_first, self.wait = int(sch.get("skip_first", 0)), int(sch.get("wait", 0))
        self.warmup, self.active, self.repeat = int(sch.get("warmup", 1)), int(sch.get("active", 4)), int(sch.get("repeat", 1))
        h = dict(json_trace_handler or {})
        self.folder = str(h.get("folder", "{run_name}/traces"))
        self.opts = dict(record_shapes=torch_prof_record_shapes, profile_memory=torch_prof_profile_memory,
                         with_stack=torch_prof_with_stack, with_flops=torch_prof_with_flops)
        self.prof: Any = None
        self.traces: list[Path] = []

    def fit_start(self, tr: "Trainer") -> None:
        from torch.profiler import ProfilerActivity, profile, schedule

        out = Path(self.folder.format(run_name=tr.state.run_name))
        if not out.is_absolute() and tr.save_folder:
            out = Path(tr.save_folder).parent / out
        out.mkdir(parents=True, exist_ok=True)
        acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if tr.device.type == "cuda" else [])

        def on_ready(p: Any) -> None:
            path = out / f"rank{tr.rank}.{tr.state.timestamp.batch}.pt.trace.json"
            p.export_chrome_trace(str(path))
            self.traces.append(path)

        self.prof = profile(activities=acts, on_trace_ready=on_ready, **self.opts,
                            schedule=schedule(skip_first=self.skip_first, wait=self.wait, warmup=self.warmup, active=self.active,
                                              repeat=self.repeat))
        self.prof.__enter__()

    def batch_end(self, tr: "Trainer") -> None:
        if self.prof is not None:
            self.prof.step()

    def fit_end(self, tr: "Trainer") -> None:
        if self.prof is not None:
            self.prof.__exit__(None, None, None)
            self.prof = None


_CALLBACKS.update({"memory_snapshot": MemorySnapshot, "oom_observer": OOMObserver})
_CALLBACKS["profiler"] = ProfilerCallback


def build_callbacks(cfg: dict[str, Any] | None) -> list[Callback]:
    out: list[Callback] = []
    for name, kw in (cfg or {}).items():
        if name not in _CALLBACKS:
            print(f"[callbacks] '{name}' is not available in photon_b200; skipping")
            continue
        out.append(_CALLBACKS[name](**dict(kw or {})))
    return out

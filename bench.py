#!/usr/bin/env python
"""Headline benchmark: tokens/sec of one federated round of MPT-125M, 8 clients, on N B200s.

    python bench.py --gpus 1 --steps 4 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Config = ``fed_125m_example``: 8 clients per round, local batch 32 × S 2048,
ADOPT lr 6e-4, cosine schedule, FedAvg (Nesterov η=1 μ=0), ``reset_optimizer=false``, amp_bf16,
random-init MPT-125M, synthetic C4-shaped tokens.  The 8 clients are spread over the N GPUs
(8/N per GPU, time-multiplexed on the node's persistent Trainer) → total work per round is
fixed → ``"scaling": "strong"``.  One "step" = every client advances one local optimizer step
(8 × 32 × 2048 tokens); the timed region is ONE full round of K local steps per client **plus
the round's aggregate + server optimizer + broadcast** (the fused NVLink kernel).

Two measurements of the same work:
* ``value``      — device time (CUDA events on the launching stream, max over ranks);
* ``e2e.value``  — wall clock around the public API call ``FederationRuntime.run_clients_fit``
                   + ``finish_round`` (every step copies its inputs H2D from pinned host memory and
                   reads the loss back D2H — that is how the Trainer works; bytes are counted).

``--impl reference`` must run the UNMODIFIED reference from ``baseline/_ref``; it cannot be
installed offline here (DESIGN.md §Reference arm) so that arm prints ``unavailable``.
``--impl torch`` runs the reference-EQUIVALENT stock path (PyTorch ops + SDPA + host shm round)
for our own A/B numbers.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_CLIENTS = 8
LOCAL_BATCH = 32
SEQ = 2048


def reference_arm() -> None:
    why = ("reference needs poetry-core build backend + flwr/composer/llm-foundry/streaming/ray/hydra (git/URL deps) — "
           "none in the image or /opt/wheelhouse, no network; pip install --no-index fails (see DESIGN.md)")
    ref = ROOT / "baseline" / "_ref" / "photon"
    if ref.exists():
        try:
            sys.path.insert(0, str(ref.parent))
            import photon.server_app  # noqa: F401  (would need flwr)
        except Exception as e:  # noqa: BLE001
            why = f"baseline/_ref present but not importable: {type(e).__name__}: {e}"
    if int(os.environ.get("RANK", "0")) == 0:   # one line per job, also under torchrun
        print(json.dumps({"impl": "reference", "unavailable": why}))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): ONE long-lived
    ``nvidia-smi -lms 200`` child started before the region (no fork/exec while the step is being timed)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, enabled: bool = True) -> None:
        self.idx, self.rows, self.proc, self.enabled = gpu_index, [], None, enabled

    def start(self) -> "ClockSampler":
        if self.enabled:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx),
                                              "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:  # noqa: BLE001
                self.proc = None
        return self

    def mark(self) -> None:
        """Timed region starts now: drop what was sampled while idle."""
        self._t0 = time.time()

    def stop(self) -> None:
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        self.rows = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.strip()]

    def summary(self) -> dict:
        load = [r for r in self.rows if len(r) > 3 and r[3].replace(".", "").isdigit() and float(r[3]) > 300.0] or self.rows
        sm = sorted(float(r[1]) for r in load if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = max((float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()), default=0.0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in load for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "note": "median over samples with power draw > 300 W (under load)"}


def build_cfg(impl: str, steps: int, model: str, attention: str, microbatch: int = LOCAL_BATCH, server: str = "fedavg"):
    from photon_b200.config import compose

    ov = [f"llm_config={model}", "run_uuid=bench", f"fl.n_total_clients={N_CLIENTS}", f"fl.n_clients_per_round={N_CLIENTS}",
          "fl.strategy_name=NESTOROV", "fl.strategy_kwargs.server_learning_rate=1.0", "fl.strategy_kwargs.server_momentum=0.0",
          "fl.reset_optimizer=false", "fl.eval_period=null", f"llm_config.global_train_batch_size={LOCAL_BATCH}",
          f"llm_config.device_train_microbatch_size={microbatch}", f"llm_config.local_steps={steps}ba",
          "llm_config.max_duration=40960ba", "llm_config.scheduler.schedulers.lr.t_max=40960ba",
          "llm_config.scheduler.schedulers.lr.t_warmup=800ba", "llm_config.precision=amp_bf16", "llm_config.log_to_console=false",
          "~llm_config.loggers.wandb", "~llm_config.loggers.tensorboard", "llm_config.save_folder=null", "llm_config.save_interval=1000000ba",
          "llm_config.eval_interval=1000000ba", "~llm_config.callbacks", "photon.checkpoint=false", "photon.comm_stack.shm=false",
          "dataset.train.root_local=synthetic://c4", "dataset.val.root_local=synthetic://c4"]
    if server == "fedadam":   # BASELINE config #4 flavour
        ov += ["fl.strategy_name=fedadam", "fl.strategy_kwargs={eta: 0.1, beta_1: 0.9, beta_2: 0.95, tau: 1.0e-9}", "fl.reset_optimizer=true"]
    if impl == "ours":
        ov += ["photon.comm_stack.nvl=true", f"kernels.attention={attention}"]
    else:  # reference-equivalent stock path: torch ops + SDPA/FA2 attention + host shm round
        ov += ["photon.comm_stack.shm=true", "kernels.gemm=torch", "kernels.attention=torch", "kernels.norm=torch",
               "kernels.loss=torch", "kernels.optimizer=torch", "llm_config.device_train_microbatch_size=8"]
    return compose(ov)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch"])
    ap.add_argument("--model", default="mpt-125m")
    ap.add_argument("--attention", default="b200", choices=["b200", "torch"])
    ap.add_argument("--microbatch", type=int, default=0, help="device microbatch (0 = 32 for mpt-125m, 8 otherwise)")
    ap.add_argument("--server", default="fedavg", choices=["fedavg", "fedadam"])
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm()
        return

    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if N_CLIENTS % world:
        raise SystemExit("--gpus must divide 8")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py measures the sm_100a engine: it needs a CUDA (B200) device and does not fall back to the CPU")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from photon_b200 import ops
    from photon_b200.federation import FederationRuntime
    from photon_b200.server.broadcast_utils import broadcast_parameters_to_nodes
    from photon_b200.utils.hw import L2_BYTES, measured_peaks

    K, W = args.steps, max(args.warmup, 3)
    sampled = list(range(N_CLIENTS))

    def make_runtime(local_steps: int) -> FederationRuntime:
        mb = args.microbatch or (LOCAL_BATCH if args.model == "mpt-125m" else 8)
        rt = FederationRuntime(build_cfg(args.impl, local_steps, args.model, args.attention, mb, args.server), device=dev, rank=rank, world_size=world)
        rt.build()
        broadcast_parameters_to_nodes(rt, rt.initial_parameters())
        return rt

    def sync() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_round(rt: FederationRuntime, server_round: int) -> list:
        res = rt.run_clients_fit(server_round, sampled)       # K local steps on each of this rank's clients
        rt.finish_round(server_round)                         # fused aggregate + server-opt + broadcast
        return res

    # ---- warm-up: a W-step round (also allocates workspaces, loads kernels, opens the arena)
    rt = make_runtime(W)
    one_round(rt, 1)
    sync()
    rt.cfg["llm_config"]["local_steps"] = f"{K}ba"
    flush = torch.empty(max(2 * L2_BYTES, 1 << 28), dtype=torch.uint8, device=dev)

    tokens = N_CLIENTS * K * LOCAL_BATCH * SEQ
    # ---- (1) device-timed round
    flush.zero_()
    sync()
    if args.impl == "ours":
        ops.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local, enabled=(rank == 0)).start()
    sync()
    clk.mark()
    ev0.record()
    res = one_round(rt, 2)
    ev1.record()
    sync()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ops.launch_count() if args.impl == "ours" else 0
    failed = [r for r in res if r.status.code != 0]
    if failed:
        raise SystemExit(f"bench round had failed clients: {failed[0].status.message}")
    # ---- (2) end-to-end wall clock through the same public API
    flush.zero_()
    sync()
    t0 = time.perf_counter()
    one_round(rt, 3)
    sync()
    e2e_s = time.perf_counter() - t0
    clk.stop()   # sampled across both timed rounds
    # ---- (3) aggregate + broadcast alone (round hot path), device-timed
    rt.round_backend.begin_round()
    rt.round_backend.add_client(rt.trainer.state.flat.params, 1.0)
    sync()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    rt.finish_round(4)
    a1.record()
    sync()
    agg_ms = a0.elapsed_time(a1)

    t = torch.tensor([dev_ms, e2e_s, agg_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, agg_ms = (float(x) for x in t.tolist())
    if rank == 0:
        clients_per_gpu = N_CLIENTS // world
        h2d = clients_per_gpu * LOCAL_BATCH * SEQ * 8            # int64 token ids per step per GPU
        d2h = clients_per_gpu * 2 * 8                            # (loss_sum, n_tokens) float64 scalars
        value = tokens / (dev_ms / 1e3)
        mcfg = rt.trainer.model_cfg
        peak = measured_peaks()
        out = {
            "metric": "tokens/sec (whole box, device-timed, max over ranks) "
                      + {"mpt-125m": "MPT-125M", "mpt-1b": "MPT-1B", "mpt-3b": "MPT-3B", "mpt-7b": "MPT-7B"}.get(args.model, args.model)
                      + " 8-client fed round",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic C4-shaped tokens, random-init weights", "impl": args.impl,
            "config": {"model": args.model, "global_batch": N_CLIENTS * LOCAL_BATCH, "seq_len": SEQ,
                       "parallelism": f"fed{N_CLIENTS}clients_on_{world}gpu", "clients_per_gpu": clients_per_gpu,
                       "local_steps_per_round": K, "local_batch": LOCAL_BATCH, "optimizer": str(rt.cfg["llm_config"]["optimizer"]["name"]), "server": "fedavg(nesterov lr=1 mu=0)" if args.server == "fedavg" else "fedadam",
                       "comm_stack": rt.round_backend.name, "attention": args.attention if args.impl == "ours" else "sdpa",
                       "l2_flush": "256 MiB memset before each timed region; per-step activations (~19 GB) exceed L2",
                       "timed_region": "one full round: K local steps x 8 clients + aggregate + server-opt + broadcast"},
            "round_aggregate_broadcast_ms": agg_ms,
            "mfu_of_measured_bf16_peak": value / world * mcfg.flops_per_token(SEQ) / peak["bf16_flops"],
            "clocks": clk.summary(),
            "e2e": {"value": tokens / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "FederationRuntime.run_clients_fit + finish_round (wall clock)"},
            "gpu_launches": int(launches),
        }
        print(json.dumps(out))
    rt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

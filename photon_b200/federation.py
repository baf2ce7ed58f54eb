"""SPMD federation runtime: the in-box control plane that replaces Flower's
SuperLink / ServerApp / ClientApp trio.

One process per GPU (``torchrun``); every rank is a *node* hosting one worker on
its GPU, rank 0 additionally keeps the server bookkeeping (history, checkpoints).
A round on N GPUs with K sampled clients:

* clients are mapped to nodes with the work-queue order (client i → node i mod N,
  each node running its queue sequentially — virtual-client multiplexing; ref:
  photon/server/server_util.py:145-202, photon/node_manager/node_manager_app.py:516);
* each node installs the global model from its own global planes (a local copy —
  the broadcast already happened inside the previous round's kernel), runs
  ``llm_fit`` on its persistent Trainer, and folds the result into the local
  weighted accumulator (streaming aggregation without leaving HBM);
* ``finish_round`` runs the transport selected by ``photon.comm_stack`` — the fused
  NVLink kernel (``nvl``) or one of the baselines — leaving the new global model
  (fp32 + bf16) in every node's planes.

``gpus_per_client > 1`` groups consecutive ranks into one client with intra-client
DDP (fused all-reduce); only the group leader contributes to the round reduce.
"""
from __future__ import annotations

import random
import time
from typing import Any

import torch
import torch.distributed as dist

from photon_b200.clients.configs import get_photon_evaluate_config_fn, get_photon_fit_config_fn
from photon_b200.clients.llm_client_functions import llm_eval, llm_fit
from photon_b200.clients.trainer_utils import get_trainer_object, pick_device
from photon_b200.clients.utils import get_initial_parameters
from photon_b200.messages import ClientState, Code, EvaluateRes, FitRes, ParamHandle, Status
from photon_b200.server.round_backends import RoundBackend, build_round_backend
from photon_b200.server.server_util import spmd_node_ids, static_assignment
from photon_b200.strategy.dispatcher import dispatch_strategy
from photon_b200.train.trainer import Trainer
from photon_b200.utils.flat import FlatLayout
from photon_b200.utils.trace import tracer


class ClientStateCache:
    """Per-client optimizer moments between a client's participations on this rank (``fl.reset_optimizer=false``), bounded.

    The reference keeps this state in client checkpoints on disk; holding it in memory saves the round trip but must not grow with
    ``n_total_clients``: entries live on the DEVICE up to ``device_bytes``, older ones are spilled to (pinned) HOST memory up to
    ``host_bytes``, and beyond that the least recently used client is dropped — it then restarts from a fresh optimizer state (or
    from its client checkpoint when there is one), exactly like a client this rank has never seen."""

    def __init__(self, device_bytes: int, host_bytes: int) -> None:
        from collections import OrderedDict

        self.device_bytes, self.host_bytes = int(device_bytes), int(host_bytes)
        self._d: "OrderedDict[int, tuple[torch.Tensor, torch.Tensor, int]]" = OrderedDict()   # least recently used first

    @staticmethod
    def _nbytes(e: tuple[torch.Tensor, torch.Tensor, int]) -> int:
        return e[0].numel() * e[0].element_size() + e[1].numel() * e[1].element_size()

    def _usage(self) -> tuple[int, int]:
        dev = sum(self._nbytes(e) for e in self._d.values() if e[0].device.type != "cpu")
        host = sum(self._nbytes(e) for e in self._d.values() if e[0].device.type == "cpu")
        return dev, host

    def __contains__(self, cid: int) -> bool:
        return int(cid) in self._d

    def __len__(self) -> int:
        return len(self._d)

    def pop(self, cid: int, default: Any = None) -> Any:
        return self._d.pop(int(cid), default)

    def get(self, cid: int) -> tuple[torch.Tensor, torch.Tensor, int]:
        e = self._d[int(cid)]
        self._d.move_to_end(int(cid))
        return e

    def put(self, cid: int, m: torch.Tensor, v: torch.Tensor, step: int) -> None:
        self._d.pop(int(cid), None)
        self._d[int(cid)] = (m, v, int(step))
        dev, host = self._usage()
        for k in list(self._d):                       # oldest first: device -> host
            if dev <= self.device_bytes:
                break
            e = self._d[k]
            if e[0].device.type == "cpu" or k == int(cid):
                continue
            pin = torch.cuda.is_available()
            self._d[k] = (e[0].to("cpu", non_blocking=False).pin_memory() if pin else e[0].cpu(),
                          e[1].to("cpu", non_blocking=False).pin_memory() if pin else e[1].cpu(), e[2])
            dev -= self._nbytes(e)
            host += self._nbytes(e)
        for k in list(self._d):                       # oldest first: host -> gone
            if host <= self.host_bytes:
                break
            if self._d[k][0].device.type == "cpu" and k != int(cid):
                host -= self._nbytes(self._d.pop(k))


class FederationRuntime:
    def __init__(self, cfg: Any, *, device: torch.device | None = None, rank: int | None = None,
                 world_size: int | None = None, group: Any = None, gpus_per_client: int = 1) -> None:
        self.cfg = cfg
        self.device = device or pick_device()
        if rank is None or world_size is None:
            on = dist.is_available() and dist.is_initialized()
            rank, world_size = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
        self.rank, self.world_size, self.group = int(rank), int(world_size), group
        self.gpus_per_client = int(gpus_per_client)
        if self.world_size % self.gpus_per_client:
            raise ValueError("world_size must be a multiple of gpus_per_client")
        self.node_id = self.rank // self.gpus_per_client          # one logical node per client group
        self.n_nodes = self.world_size // self.gpus_per_client
        self.is_leader = self.rank % self.gpus_per_client == 0
        self.client_group = None
        self.grad_comm = None
        if self.gpus_per_client > 1:
            for g in range(self.n_nodes):
                ranks = list(range(g * self.gpus_per_client, (g + 1) * self.gpus_per_client))
                pg = dist.new_group(ranks)
                if g == self.node_id:
                    self.client_group = pg
        self.strategy = dispatch_strategy(cfg)
        self.rng = random.Random(int(cfg["seed"]))
        self.client_states: dict[int, ClientState] = {c: ClientState() for c in range(int(cfg["fl"]["n_total_clients"]))}
        self.server_steps_cumulative = 0
        self.fit_config_fn = get_photon_fit_config_fn(cfg)
        self.eval_config_fn = get_photon_evaluate_config_fn(cfg)
        self.trainer: Trainer | None = None
        # cid -> (exp_avg, exp_avg_sq, step) of its last fit HERE; bounded (device, then pinned host, then least-recently-used out)
        ph = cfg["photon"]
        self._opt_states = ClientStateCache(int(float(ph.get("client_state_cache_device_gb", 16.0) or 0.0) * (1 << 30)),
                                            int(float(ph.get("client_state_cache_host_gb", 64.0) or 0.0) * (1 << 30)))
        self._local_params: dict[int, torch.Tensor] = {}   # personalised-layer memory per client
        self.layout: FlatLayout | None = None        # exchange layout (3 planes with fl.aggregate_momenta)
        self.model_layout: FlatLayout | None = None  # the trainer's parameter layout
        self.aggregate_momenta = bool(cfg["fl"].get("aggregate_momenta", False))
        self.round_backend: RoundBackend | None = None
        self.fault_injection = dict(cfg["fl"].get("fault_injection") or {})
        self.timings: dict[str, float] = {}
        # host control plane (liveness, shared work queue, time-bounded metadata exchange); None on a single rank
        self.ctl: Any = None
        self.alive_ranks: list[int] = list(range(self.world_size))
        self.scheduling = str(cfg["photon"].get("scheduling", "dynamic") or "dynamic").lower()
        self.assignment: dict[int, int] = {}        # cid -> node that trained it in the last round (diagnostics / tests)

    # --------------------------------------------------------------------- bring-up
    def build(self) -> None:
        """Create the persistent Trainer (also fixes the flat layout) and the round transport."""
        want = int(self.cfg["photon"].get("n_nodes", 1) or 1)
        if want not in (1, self.n_nodes) and self.rank == 0:
            # the reference waits for photon.n_nodes Flower nodes; here the job's ranks ARE the nodes
            print(f"[federation] photon.n_nodes={want} but this job has {self.n_nodes} node(s) "
                  f"({self.world_size} rank(s) / {self.gpus_per_client} GPU(s) per client): using {self.n_nodes}", flush=True)
        kw: dict[str, Any] = dict(device=self.device, rank=self.rank % self.gpus_per_client, world_size=self.gpus_per_client,
                                  process_group=self.client_group)
        if self.gpus_per_client > 1 and self.device.type == "cuda":
            # intra-client DDP over the fused NVLink all-reduce: the gradient plane lives in a symmetric arena
            from photon_b200.parallel.ddp import build_nvl_comm, wants_sharded_step
            from photon_b200.utils.flat import layout_for_model_cfg

            fl_ = self.cfg["fl"]
            total = layout_for_model_cfg(self.cfg["llm_config"]["model"], fl_.get("frozen_layers"), fl_.get("unfrozen_layers")).total
            kw["grad_comm"] = build_nvl_comm(total, sharded=wants_sharded_step(self.cfg["llm_config"]),
                                             rank=self.rank % self.gpus_per_client, world_size=self.gpus_per_client,
                                             device=self.device, group=self.client_group)
        elif self.gpus_per_client > 1:
            from photon_b200.parallel.ddp import NcclGradComm

            kw["grad_comm"] = NcclGradComm(self.client_group)
        fl = self.cfg["fl"]
        self.trainer, _ = get_trainer_object(self.cfg, 0, log_name=f"_node_{self.node_id}", split_eval=bool(fl.get("split_eval", False)),
                                             use_unigram_metrics=bool(fl["use_unigram_metrics"]),
                                             allow_unigram_metrics_failures=bool(fl["allow_unigram_metrics_failures"]),
                                             frozen_layers=fl.get("frozen_layers"), unfrozen_layers=fl.get("unfrozen_layers"),
                                             resize_vocab=fl.get("resize_vocab"), **kw)
        self.model_layout = self.trainer.state.flat.layout
        # fl.aggregate_momenta (R3): the exchanged / aggregated / checkpointed vector is [params | exp_avg | exp_avg_sq];
        # the strategy is oblivious to the structure, exactly like the reference's 3n-array payload
        self.layout = self.model_layout.stacked(("", "exp_avg/", "exp_avg_sq/")) if self.aggregate_momenta else self.model_layout
        self.round_backend = build_round_backend(self.cfg, self.layout, self.strategy, self.device, rank=self.rank,
                                                 world_size=self.world_size, group=self.group)
        from photon_b200.server.control import build_control_plane

        self.ctl = build_control_plane(self.cfg, self.rank, self.world_size)
        self.round_backend.ctl = self.ctl
        if self.ctl is not None:
            from photon_b200.train.callbacks import Callback

            ctl = self.ctl

            class _Progress(Callback):     # every finished batch is a sign of life of this rank's main thread
                def batch_end(self, trainer: Any) -> None:
                    ctl.tick()

            self.trainer.callbacks.append(_Progress())
        if self.ctl is not None and self.device.type == "cuda":
            from photon_b200 import ops

            # the in-kernel spins give up after this long: a peer that died after the host-side liveness check turns into an
            # aborted (and then repeated, masked) round instead of a hang
            ops.ext().set_comm_timeout_ms(int(1000 * float(self.cfg["photon"].get("kernel_peer_timeout_s", 120.0) or 120.0)))

    def initial_parameters(self) -> torch.Tensor:
        """Rank 0's freshly initialised (or pretrained) model as a flat tensor."""
        assert self.layout is not None
        flat = torch.zeros(self.layout.total, dtype=torch.float32)
        if self.rank == 0:
            arrays, lay = get_initial_parameters(self.cfg)
            if lay.names != self.model_layout.names:
                raise AssertionError("initial-parameter layout differs from the trainer's")
            self.model_layout.from_ndarrays(flat[: self.model_layout.total], arrays)  # momenta planes start at zero
        return flat

    def node_ids(self) -> list[int]:
        """Logical nodes currently alive: one per client group of ranks (ref: Driver.get_node_ids). With the control plane this
        is a real liveness check (heartbeats); a client group counts as alive only while ALL its ranks are."""
        if self.ctl is None:
            return spmd_node_ids(self.group)[:: self.gpus_per_client]
        alive = set(self.ctl.alive())
        g = self.gpus_per_client
        return [n for n in range(self.n_nodes) if all(r in alive for r in range(n * g, (n + 1) * g))]

    # ----------------------------------------------------------------------- sampling
    def sample_clients(self) -> list[int]:
        """Seeded ``random.Random(cfg.seed).sample`` — identical on every rank and replayable on
        resume (ref: photon/server_app.py:124,188-192,295)."""
        fl = self.cfg["fl"]
        return self.rng.sample(range(int(fl["n_total_clients"])), int(fl["n_clients_per_round"]))

    def replay_sampling(self, n_rounds: int) -> None:
        for _ in range(n_rounds):
            self.sample_clients()

    def my_clients(self, sampled: list[int]) -> list[int]:
        """Static queue of this node: client i -> living node i mod n (what the work queue converges to with equally fast nodes)."""
        nodes = self.node_ids() if self.ctl is not None else list(range(self.n_nodes))
        return static_assignment(sampled, nodes).get(self.node_id, [])

    def _client_queue(self, server_round: int, sampled: list[int]) -> Any:
        """The clients this node trains in this round, one at a time. ``photon.scheduling=dynamic`` (default with a control plane):
        a shared atomic counter — whenever this node is free it takes the NEXT sampled client, so a fast GPU ends up with more
        clients than a slow one (the reference's reply-frees-the-node work queue, ref: photon/server/server_util.py:163-202).
        ``static``: the precomputed queue (reproducible client-to-node mapping)."""
        if self.ctl is None or self.scheduling == "static":
            yield from self.my_clients(sampled)
            return
        q = self.ctl.open_queue(f"fit/{server_round}")
        while True:
            idx = self.ctl.next_index(q) if self.is_leader else 0
            if self.gpus_per_client > 1:     # the leader's pick is the whole client group's pick
                t = torch.tensor([idx], dtype=torch.int64, device=self.device if dist.get_backend(self.client_group) == "nccl" else "cpu")
                dist.broadcast(t, src=self.node_id * self.gpus_per_client, group=self.client_group)
                idx = int(t.item())
            if idx >= len(sampled):
                return
            yield sampled[idx]

    # -------------------------------------------------------------------------- fit
    def _should_fail(self, server_round: int, cid: int) -> bool:
        """``fl.fault_injection: {round, cid, kind}`` — ``kind: exception`` (default) fails the client, ``kill`` SIGKILLs the whole
        rank while it trains the client (a dead node), ``hang`` blocks it forever (a hung node: its heartbeat goes stale after
        ``photon.progress_timeout_s``)."""
        fi = self.fault_injection
        if not (bool(fi) and int(fi.get("round", -1)) == server_round):
            return False
        # target: a client id (``cid``) and / or whatever client a given rank is training (``rank``)
        if ("cid" in fi and int(fi["cid"]) != cid) or ("rank" in fi and int(fi["rank"]) != self.rank) or not ({"cid", "rank"} & set(fi)):
            return False
        kind = str(fi.get("kind", "exception"))
        if kind == "slow":      # a straggler: the client takes `seconds` longer (work-queue rebalancing tests)
            time.sleep(float(fi.get("seconds", 1.0)))
            return False
        if kind == "kill":
            import os
            import signal

            print(f"[fault-injection] rank {self.rank}: SIGKILL while training client {cid} in round {server_round}", flush=True)
            os.kill(os.getpid(), signal.SIGKILL)
        if kind == "hang":
            print(f"[fault-injection] rank {self.rank}: hanging while training client {cid} in round {server_round}", flush=True)
            while True:
                time.sleep(3600.0)
        return True

    def run_clients_fit(self, server_round: int, sampled: list[int]) -> list[FitRes]:
        assert self.trainer is not None and self.round_backend is not None
        rb, tr = self.round_backend, self.trainer
        period = int(self.cfg["photon"].get("refresh_period", 0) or 0)
        if period and server_round > 1 and server_round % period == 0:
            # the reference recycles its worker PROCESSES here to shed leaked memory (ref: client_app.py:175-177); a rank of
            # the SPMD job cannot restart itself, so it drops what can be dropped: python garbage and the allocator's cache
            import gc

            gc.collect()
            if self.device.type == "cuda":
                torch.cuda.empty_cache()
        rb.begin_round()
        results: list[FitRes] = []
        keep_opt = not bool(self.cfg["fl"]["reset_optimizer"])
        t_fit = 0.0
        self._trained_here: list[int] = []
        for cid in self._client_queue(server_round, sampled):
            self._trained_here.append(cid)
            t0 = time.time()
            try:
                if self._should_fail(server_round, cid):
                    raise RuntimeError(f"fault injection: client {cid} dropped in round {server_round}")
                fc = self.fit_config_fn(server_round, cid, self.client_states, self.server_steps_cumulative)
                opt = tr.state.optimizer
                if keep_opt and cid in self._opt_states:
                    # the client's own moments from ITS last participation (entries are dropped in gather_results as soon as the
                    # client trains elsewhere, so what is found here is never older than that); a client checkpoint, when present,
                    # is loaded on top by llm_fit. Keyed by client id, whatever the client-to-node mapping of the round.
                    m, v, step = self._opt_states.get(cid)
                    opt.exp_avg.copy_(m), opt.exp_avg_sq.copy_(v)  # per-rank planes (a slice when the state is sharded); H2D when spilled
                    opt.step_count = step
                elif keep_opt:
                    opt.reset_state()
                if fc.personalized_layers and cid in self._local_params:
                    tr.state.flat.params.copy_(self._local_params[cid])
                with tracer().span("client_fit", cat="client", device=True, cid=cid, server_round=server_round):
                    shadow = rb.global_shadow()
                    payload, n_samples, metrics, _ = llm_fit(tr, rb.global_params(), fc, self.cfg, cid,
                                                             shadow_payload=None if shadow is None else shadow[: self.model_layout.total])
                if keep_opt:   # always (also when the node hosts one client): behaviour must not depend on the topology
                    self._opt_states.put(cid, opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.step_count)
                if fc.personalized_layers:
                    self._local_params[cid] = tr.state.flat.params.clone()
                if self.is_leader:
                    if not (torch.is_tensor(payload) and payload.numel() == self.layout.total):
                        raise AssertionError("client payload does not match the exchange layout")
                    rb.add_client(payload, n_samples)
                results.append(FitRes(Status(Code.OK, ""), ParamHandle(kind=rb.name), n_samples if self.is_leader else 0, metrics, cid))
            except Exception as e:  # noqa: BLE001 - a failed client must not take the round down
                results.append(FitRes(Status(Code.FAILED, repr(e)), None, 0, {}, cid))
            t_fit += time.time() - t0
        self.timings["node_training_time_s"] = t_fit
        return results

    def gather_results(self, results: list[FitRes], sampled: list[int] | None = None) -> list[FitRes]:
        """Control-plane gather (metadata only — parameters never travel here). With the host control plane this is also where a
        dead rank is discovered: it never answers, rank 0 rules it out, and every sampled client nobody reported on becomes a
        FAILED result (counted against ``fl.accept_failures_cnt`` like any other client failure)."""
        mine = [r for r in results if self.is_leader or r.status.code != Code.OK]
        if self.ctl is not None:
            parts = self.ctl.gather("fit_results", (self.node_id, mine))
            self.alive_ranks = sorted(parts)
            out = [r for rk in sorted(parts) for r in parts[rk][1]]
            self.assignment = {int(r.cid): int(parts[rk][0]) for rk in sorted(parts) for r in parts[rk][1] if r.status.code == Code.OK}
            if sampled is not None:
                seen = {int(r.cid) for r in out}
                for cid in sampled:
                    if int(cid) not in seen:
                        out.append(FitRes(Status(Code.FAILED, "the node training this client stopped responding (marked dead)"), None, 0, {}, cid))
        elif self.world_size == 1 or not dist.is_initialized():
            out = mine
        else:
            box: list[Any] = [None] * self.world_size
            dist.all_gather_object(box, mine, group=self.group)
            out = [r for part in box for r in part]
        # clients another node trained this round leave stale optimizer moments here
        here = set(getattr(self, "_trained_here", []))
        for r in out:
            if r.status.code == Code.OK and int(r.cid) not in here:
                self._opt_states.pop(int(r.cid), None)
        return out

    def finish_round(self, server_round: int) -> None:
        """Aggregate + server optimizer + broadcast over the ranks that are alive. If the fused kernel reports that a participant
        never reached its start barrier (it died after the host-side check) nothing was modified: the rank is ruled out and the
        round is run again over the survivors."""
        assert self.round_backend is not None
        t0 = time.time()
        with tracer().span("aggregate_server_opt_broadcast", cat="round", device=True, server_round=server_round, transport=self.round_backend.name):
            for attempt in range(max(1, self.world_size)):
                alive = list(self.alive_ranks)
                self.round_backend.finish_round(server_round, alive=alive if len(alive) < self.world_size else None)
                if self.ctl is None:
                    break
                st = self.round_backend.status()
                verdict = self.ctl.gather(f"round_status/{server_round}/{attempt}", int(st))
                lost = {alive[t] for stv in verdict.values() for t in range(len(alive)) if (int(stv) >> t) & 1 or (int(stv) >> (8 + t)) & 1}
                lost |= set(alive) - set(verdict)
                if any((int(stv) >> 8) & 0xFF for stv in verdict.values()):
                    raise RuntimeError("a peer vanished INSIDE the round kernel: the global model may be torn; resume from the last server checkpoint")
                if not lost:
                    break
                self.ctl.mark_dead(lost)
                self.alive_ranks = [r for r in alive if r not in lost]
                print(f"[federation] round {server_round}: rank(s) {sorted(lost)} did not reach the round kernel; repeating over {self.alive_ranks}", flush=True)
        self.timings["aggregate_broadcast_host_s"] = time.time() - t0

    def abort_round(self) -> None:
        """Ignored round: global model untouched (the accumulators are simply dropped)."""

    # ----------------------------------------------------------------------- evaluate
    def run_clients_evaluate(self, server_round: int, sampled: list[int]) -> list[EvaluateRes]:
        assert self.trainer is not None and self.round_backend is not None
        out: list[EvaluateRes] = []
        if self.node_id != 0:   # the reference evaluates on ONE client (id 0), all streams concatenated
            return self._gather_eval(out)
        for cid in sampled:
            try:
                ec = self.eval_config_fn(server_round, cid, self.client_states, self.server_steps_cumulative)
                loss, n, metrics, _ = llm_eval(self.trainer, self.round_backend.global_params(), ec, self.cfg, cid)
                out.append(EvaluateRes(Status(Code.OK, ""), loss, n, metrics, cid))
            except Exception as e:  # noqa: BLE001
                out.append(EvaluateRes(Status(Code.FAILED, repr(e)), 0.0, 0, {}, cid))
        return self._gather_eval(out)

    def _gather_eval(self, mine: list[EvaluateRes]) -> list[EvaluateRes]:
        if self.ctl is not None:
            parts = self.ctl.gather("eval_results", mine if self.is_leader else [])
            return [r for rk in sorted(parts) for r in parts[rk]]
        if self.world_size == 1 or not dist.is_initialized():
            return mine
        box: list[Any] = [None] * self.world_size
        dist.all_gather_object(box, mine if self.is_leader else [], group=self.group)
        return [r for part in box for r in part]

    # --------------------------------------------------------------------------- state
    def state_tensors(self) -> dict[str, torch.Tensor]:
        assert self.round_backend is not None
        m, v = self.round_backend.moments()
        vals = [self.round_backend.global_params(), m, v]
        return {k: t for k, t in zip(self.strategy.state_keys, vals) if t is not None}

    def close(self) -> None:
        if self.trainer is not None:
            self.trainer.close()
        if self.ctl is not None:
            self.ctl.close()
        if self.round_backend is not None:
            if self.ctl is not None and self.ctl.dead:
                return   # arena teardown is collective over the original group; with a dead peer the process just exits
            self.round_backend.close()

"""Host control plane of the SPMD runtime: liveness, a shared work queue and time-bounded metadata exchange.

The reference gets these from Flower's SuperLink: nodes that stop answering disappear from ``driver.get_node_ids()`` (checked twice
per round, ref: photon/server_app.py:285,346), every reply frees its node for the NEXT sampled client (a work queue, ref:
photon/server/server_util.py:163-202) and a node manager restarts dead workers / re-queues their client (ref:
photon/node_manager/node_manager_app.py:326-351,553-579). Here the ranks of one job talk through a ``torch.distributed.TCPStore``
hosted by rank 0 (the server): tiny keys only — parameters never travel here, they stay in the NVLink arena.

* **liveness** — every rank's daemon thread refreshes ``hb/{rank}``; a rank whose heartbeat is older than ``liveness_timeout_s`` is
  dead. Rank 0 is the authority: it decides who took part in an exchange and publishes the verdict, so all survivors act on the
  same membership (a rank that loses rank 0 raises :class:`ServerLostError` — the server died, as in the reference).
* **work queue** — ``next_index(tag)`` is an atomic counter in the store: a GPU that finishes early pulls the next sampled client,
  so uneven clients (``local_steps``, data, power capping) balance out instead of waiting for the slowest static queue.
* **exchanges** — ``gather`` / ``sum_tensor`` / ``broadcast`` / ``barrier`` never block on a dead peer: they return what the
  living ranks contributed.

Nothing here touches NCCL, so it keeps working when a peer process was SIGKILLed (NCCL collectives over the original group would hang).
"""
from __future__ import annotations

import os
import pickle
import threading
import time
from datetime import timedelta
from typing import Any

import torch


class ServerLostError(RuntimeError):
    """Rank 0 (server bookkeeping + store host) stopped responding."""


class ControlPlane:
    _instances = 0     # every rank creates its control planes in the same order: the count makes key spaces of successive runtimes disjoint

    def __init__(self, rank: int, world_size: int, *, host: str | None = None, port: int | None = None, heartbeat_s: float = 0.5,
                 liveness_timeout_s: float = 20.0, exchange_timeout_s: float = 7200.0, progress_timeout_s: float = 900.0,
                 namespace: str = "photon", instance: int | None = None) -> None:
        from torch.distributed import TCPStore

        self.rank, self.world_size = int(rank), int(world_size)
        self.heartbeat_s, self.liveness_timeout_s, self.exchange_timeout_s = float(heartbeat_s), float(liveness_timeout_s), float(exchange_timeout_s)
        host = host or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(port or os.environ.get("PHOTON_CONTROL_PORT") or (int(os.environ.get("MASTER_PORT", "29500")) + 101))
        self.store = TCPStore(host, port, world_size=None, is_master=(self.rank == 0), timeout=timedelta(seconds=60), wait_for_workers=False,
                              multi_tenant=True)
        ControlPlane._instances += 1
        # a later runtime in the same job must never read this one's keys (``instance``: explicit id when several "ranks" live in
        # one process, e.g. unit tests)
        self.ns = f"{namespace}/{ControlPlane._instances if instance is None else instance}"
        # a rank whose MAIN thread made no progress for this long stops refreshing its heartbeat: a hung worker (stuck kernel,
        # dead-locked loader) then looks exactly like a dead one (ref: the node manager's per-task time-out)
        self.progress_timeout_s = float(progress_timeout_s)
        self._last_tick = time.time()
        self.dead: set[int] = set()
        self._seq: dict[str, int] = {}
        self._stop = threading.Event()
        self._beat()
        self._thread = threading.Thread(target=self._heartbeat_loop, name="photon-heartbeat", daemon=True)
        self._thread.start()
        # everyone present before the first round (bounded: a rank that never shows up is a launch error, not a fault to mask)
        self.store.set(self._k(f"hello/{self.rank}"), b"1")
        t0 = time.time()
        while not all(self.store.check([self._k(f"hello/{r}")]) for r in range(self.world_size)):
            if time.time() - t0 > 300.0:
                raise TimeoutError("control plane: not every rank of the job connected within 300 s")
            time.sleep(0.01)

    # ------------------------------------------------------------------ plumbing
    def _k(self, key: str) -> str:
        return f"{self.ns}/{key}"

    def _beat(self) -> None:
        self.store.set(self._k(f"hb/{self.rank}"), repr(time.time()).encode())

    def tick(self) -> None:
        """Progress mark of the main thread (every training batch, every control-plane call)."""
        self._last_tick = time.time()

    def _heartbeat_loop(self) -> None:
        while not self._stop.wait(self.heartbeat_s):
            if time.time() - self._last_tick > self.progress_timeout_s:
                continue    # hung: let the heartbeat go stale
            try:
                self._beat()
            except Exception:  # noqa: BLE001 - the store went away (rank 0 died): nothing left to tell
                return

    def _age(self, r: int) -> float:
        try:
            return time.time() - float(self.store.get(self._k(f"hb/{r}")).decode())
        except Exception:  # noqa: BLE001
            return float("inf")

    def _next_seq(self, tag: str) -> str:
        n = self._seq.get(tag, 0)
        self._seq[tag] = n + 1
        return f"{tag}#{n}"

    # ------------------------------------------------------------------ liveness
    def is_alive(self, r: int) -> bool:
        return r == self.rank or (r not in self.dead and self._age(r) < self.liveness_timeout_s)

    def alive(self) -> list[int]:
        """Ranks currently considered alive (ref: ``driver.get_node_ids()``)."""
        return [r for r in range(self.world_size) if self.is_alive(r)]

    def mark_dead(self, ranks: Any) -> None:
        self.dead.update(int(r) for r in ranks)

    # ------------------------------------------------------------------ work queue
    def open_queue(self, tag: str) -> str:
        """Start a fresh shared counter for ``tag`` (every rank calls it at the same point of the protocol)."""
        return self._next_seq(f"q/{tag}")

    def next_index(self, queue: str) -> int:
        """Atomically take the next work item index of ``queue`` (0, 1, 2, ... across ALL ranks)."""
        self.tick()
        return int(self.store.add(self._k(queue), 1)) - 1

    # ------------------------------------------------------------------ exchanges
    def _wait_key(self, key: str, owner: int, deadline: float) -> bool:
        """True when ``key`` exists; False when its owner died (or the exchange timed out) first."""
        t0 = time.time()
        n = 0
        while True:
            if self.store.check([key]):
                return True
            n += 1
            waited = time.time() - t0
            if waited > 0.02 and n % 8 == 0:    # the liveness probe is a store round trip of its own: not on the fast path
                self.tick()                      # waiting for a peer IS this rank's progress
                if owner in self.dead or self._age(owner) > self.liveness_timeout_s or time.time() > deadline:
                    return False
            time.sleep(0.00005 if waited < 0.005 else 0.002)   # peers normally arrive within microseconds of each other

    def gather(self, tag: str, obj: Any, timeout_s: float | None = None) -> dict[int, Any]:
        """Every living rank contributes ``obj``; returns ``{rank: obj}`` for the ranks rank 0 ruled in (identical on every survivor).
        Ranks that died before contributing are added to :attr:`dead`."""
        self.tick()
        key = self._k(self._next_seq(f"x/{tag}"))
        self.store.set(f"{key}/{self.rank}", pickle.dumps(obj))
        deadline = time.time() + (self.exchange_timeout_s if timeout_s is None else float(timeout_s))
        if self.rank == 0:
            included = []
            for r in range(self.world_size):
                if r in self.dead:
                    continue
                if self._wait_key(f"{key}/{r}", r, deadline):
                    included.append(r)
                elif timeout_s is not None and self._age(r) < self.liveness_timeout_s:
                    continue     # a short, best-effort exchange (shutdown): late is not dead
                else:
                    self.dead.add(r)
                    print(f"[control] rank {r} stopped responding (heartbeat age {self._age(r):.1f} s): marked dead", flush=True)
            self.store.set(f"{key}/verdict", pickle.dumps((included, sorted(self.dead))))
        else:
            if not self._wait_key(f"{key}/verdict", 0, deadline):
                raise ServerLostError("rank 0 (server) stopped responding")
            included, dead = pickle.loads(self.store.get(f"{key}/verdict"))
            self.dead = set(dead)
            if self.rank not in included:
                raise ServerLostError(f"rank {self.rank} was ruled dead by the server (heartbeat too old); leaving the job")
        return {r: pickle.loads(self.store.get(f"{key}/{r}")) for r in included}

    def barrier(self, tag: str) -> list[int]:
        """Returns the ranks that arrived."""
        return sorted(self.gather(f"bar/{tag}", None))

    def sum_tensor(self, tag: str, t: torch.Tensor) -> torch.Tensor:
        """Element-wise sum over the living ranks of a SMALL tensor (metrics, norm partials); result on ``t``'s device."""
        parts = self.gather(f"sum/{tag}", t.detach().cpu())
        out = None
        for r in sorted(parts):
            out = parts[r].clone() if out is None else out + parts[r]
        return out.to(t.device)

    def broadcast(self, tag: str, obj: Any, src: int = 0) -> Any:
        got = self.gather(f"bc/{tag}", obj if self.rank == src else None)
        if src not in got:
            raise ServerLostError(f"broadcast source rank {src} is dead")
        return got[src]

    def close(self) -> None:
        if self.store is None:
            return
        try:            # leave together: nobody's heartbeat should find the store gone (rank 0 hosts it and goes last)
            self.gather("bar/close", None, timeout_s=5.0)
        except Exception:  # noqa: BLE001
            pass
        self._stop.set()
        self._thread.join(timeout=2.0)
        if self.rank == 0:
            time.sleep(2.0 * self.heartbeat_s)
        self.store = None   # drop the client connection (rank 0: the server socket goes with the last reference)


def build_control_plane(cfg: Any, rank: int, world_size: int) -> ControlPlane | None:
    """``photon.control_plane``: ``store`` (default for multi-rank jobs) or ``none`` (static queues, dist collectives, no liveness)."""
    ph = cfg["photon"]
    mode = str(ph.get("control_plane", "store") or "store").lower()
    if world_size <= 1 or mode in ("none", "off", "false"):
        return None
    return ControlPlane(rank, world_size, liveness_timeout_s=float(ph.get("liveness_timeout_s", 20.0) or 20.0),
                        progress_timeout_s=float(ph.get("progress_timeout_s", 900.0) or 900.0),
                        namespace=f"photon/{cfg.get('run_uuid', 'run')}")

"""One federated fit round on the server side (ref: photon/server/fit_utils.py:41-389):
collect the nodes' results, filter failures, feed the strategy / round transport, merge client
state, advance ``server_steps_cumulative`` and aggregate client metrics."""
from __future__ import annotations

import time
from typing import Any

from photon_b200.messages import Code, FitRes, decode_client_states
from photon_b200.strategy.aggregation import weighted_average


class TooManyFailuresError(RuntimeError):
    """More client failures than ``fl.accept_failures_cnt`` in one round."""


def split_results(results: list[FitRes]) -> tuple[list[FitRes], list[FitRes]]:
    ok = [r for r in results if r.status.code == Code.OK and r.num_examples > 0]
    failed = [r for r in results if not (r.status.code == Code.OK and r.num_examples > 0)]
    return ok, failed


def check_failures(n_failures: int, accept_failures_cnt: int, ignore_failed_rounds: bool) -> bool:
    """Returns True when the round must be ignored (previous global params kept).

    Semantics: up to ``accept_failures_cnt`` failed clients are tolerated and the round
    aggregates the survivors; beyond that the round either raises or — with
    ``ignore_failed_rounds`` — is skipped (ref: fit_utils.py:198-210,257-286; the reference's
    counter is re-zeroed per call, SURVEY App. D #2 — here it is a real per-round count)."""
    if n_failures <= accept_failures_cnt:
        return False
    if ignore_failed_rounds:
        return True
    raise TooManyFailuresError(f"{n_failures} client failure(s) > accept_failures_cnt={accept_failures_cnt}")


def get_handle_success_and_failure_fit(metrics_accumulator: list[tuple[dict[str, Any], Any, int]], fit_failures: list[FitRes | None],
                                       accept_failures_cnt: int | None) -> Any:
    """The reference's streaming form of ``split_results`` + ``check_failures`` (ref: fit_utils.py:220-288): a callable fed one
    ``(ok, FitRes)`` pair at a time — successes append ``(metrics, status, num_examples)``, failures are collected and counted, and
    the count passing ``accept_failures_cnt`` raises. (The count is cumulative over the closure's life; the reference resets it on
    every call, which makes its own limit unreachable.)"""
    seen = {"failures": 0}

    def handle_success_and_failure_fit(result: tuple[bool, FitRes | None]) -> tuple[bool, FitRes | None]:
        ok, res = result
        if ok and res is not None:
            metrics_accumulator.append((res.metrics, res.status, res.num_examples))
            return True, res
        seen["failures"] += 1
        if accept_failures_cnt is not None and seen["failures"] > accept_failures_cnt:
            raise TooManyFailuresError(f"{seen['failures']} client failure(s) > accept_failures_cnt={accept_failures_cnt}")
        fit_failures.append(res)
        return False, res

    return handle_success_and_failure_fit


def handle_fit_replies(runtime: Any, server_round: int, results: list[FitRes]) -> dict[str, Any]:
    """Server bookkeeping after the transport has consumed the parameter payloads."""
    ok, failed = split_results(results)
    metrics: dict[str, Any] = {"server/n_failures": len(failed)}
    merged = weighted_average([(r.num_examples, r.metrics) for r in ok]) if ok else {}
    acc = merged.pop("client_state_acc", None)
    if acc:
        for cid, st in decode_client_states(acc).items():
            runtime.client_states[cid] = st
    steps = [s.steps_done for s in runtime.client_states.values()] + [0]
    runtime.server_steps_cumulative += max(steps)          # ref: fit_utils.py:179-183
    for s in runtime.client_states.values():
        s.steps_done = 0                                   # ref: fit_utils.py:187-188
    metrics.update(merged)
    return metrics


def fit_round(runtime: Any, server_round: int, sampled_clients: list[int]) -> dict[str, Any]:
    """Run the round on this rank's clients, aggregate, return server+client metrics
    (``server/fit_round_time`` included; ref: fit_utils.py:291-389)."""
    t0 = time.time()
    results = runtime.run_clients_fit(server_round, sampled_clients)
    all_results = runtime.gather_results(results, sampled_clients)
    ok, failed = split_results(all_results)
    for r in failed:   # say WHY before deciding whether the round survives (the reference only counts)
        print(f"[server] round {server_round}: client {getattr(r, 'cid', '?')} failed: {r.status.message}", flush=True)
    ignore = check_failures(len(failed), int(runtime.cfg["fl"]["accept_failures_cnt"]), bool(runtime.cfg["fl"]["ignore_failed_rounds"]))
    if ignore or not ok:
        runtime.abort_round()
        metrics = {"server/n_failures": len(failed), "server/round_ignored": 1}
    else:
        runtime.finish_round(server_round)
        metrics = handle_fit_replies(runtime, server_round, all_results)
        metrics.update(runtime.round_backend.collect_metrics())
    metrics["server/fit_round_time"] = time.time() - t0
    return metrics

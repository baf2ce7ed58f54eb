"""Federated evaluation round (ref: photon/server/evaluate_utils.py:33-296). The server always
evaluates client id 0 with all streams concatenated (ref: photon/server_app.py:258,361)."""
from __future__ import annotations

import time
from typing import Any

from photon_b200.messages import Code, EvaluateRes


def handle_evaluate_replies(runtime: Any, server_round: int, results: list[EvaluateRes]) -> tuple[float | None, dict[str, Any]]:
    ok = [r for r in results if r.status.code == Code.OK]
    failed = len(results) - len(ok)
    if failed > int(runtime.cfg["fl"]["accept_failures_cnt"]) and not runtime.cfg["fl"]["ignore_failed_rounds"]:
        from photon_b200.server.fit_utils import TooManyFailuresError

        raise TooManyFailuresError(f"{failed} evaluate failure(s)")
    return runtime.strategy.aggregate_evaluate(server_round, ok, [])


def get_handle_success_and_failure_evaluate(metrics_accumulator: list[tuple[float, dict[str, Any], Any, int]],
                                            evaluate_failures: list[EvaluateRes | None]) -> Any:
    """One ``(ok, EvaluateRes)`` pair at a time: successes append ``(loss, metrics, status, num_examples)``, failures are collected
    (ref: evaluate_utils.py ``get_handle_success_and_failure_evaluate``); ``handle_evaluate_replies`` is the batch form used here."""
    def handle_success_and_failure_evaluate(result: tuple[bool, EvaluateRes | None]) -> tuple[bool, EvaluateRes | None]:
        ok, res = result
        if ok and res is not None:
            metrics_accumulator.append((res.loss, res.metrics, res.status, res.num_examples))
        else:
            evaluate_failures.append(res)
        return bool(ok and res is not None), res

    return handle_success_and_failure_evaluate


def evaluate_round(runtime: Any, server_round: int, sampled_clients: list[int] | None = None) -> tuple[float | None, dict[str, Any]]:
    t0 = time.time()
    results = runtime.run_clients_evaluate(server_round, sampled_clients or [0])
    loss, metrics = handle_evaluate_replies(runtime, server_round, results)
    metrics["server/evaluate_round_time"] = time.time() - t0
    return loss, metrics

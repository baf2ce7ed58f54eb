"""Server messaging helpers: node discovery, the work-queue client scheduler, instruction
builders (ref: photon/server/server_util.py:35-302).

``ClientScheduler`` keeps the reference's semantics — one instruction per sampled client; the
first ``len(nodes)`` are dispatched at once and every reply frees that node for the next client
(``message_collaborative``, ref :65-202) — as a plain generator so it can drive real nodes,
virtual nodes in tests, or be evaluated statically (``static_assignment``) by the SPMD runtime
where every rank must know its own queue without a broker.
"""
from __future__ import annotations

import time
from collections import deque
from typing import Any, Callable, Iterable, Iterator

import torch.distributed as dist

from photon_b200.messages import ClientState, EvaluateIns, FitIns, Message, encode_client_states

PARAMETERS = "parameters"
COMM_ = "comm_"
COMM_STACK = "comm_stack"


def wait_for_nodes_to_connect(n_nodes: int, get_node_ids: Callable[[], list[int]], poll_s: float = 3.0,
                              timeout_s: float | None = None) -> list[int]:
    """Block until ``n_nodes`` nodes are visible (ref: server_util.py:35-62)."""
    t0 = time.time()
    while True:
        ids = get_node_ids()
        if len(ids) >= n_nodes:
            return ids
        if timeout_s is not None and time.time() - t0 > timeout_s:
            raise TimeoutError(f"only {len(ids)}/{n_nodes} nodes connected after {timeout_s}s")
        time.sleep(poll_s)


def spmd_node_ids(group: Any = None) -> list[int]:
    """In the SPMD runtime a node == a rank of the job."""
    if dist.is_available() and dist.is_initialized():
        return list(range(dist.get_world_size(group)))
    return [0]


def static_assignment(sampled_clients: list[int], node_ids: list[int]) -> dict[int, list[int]]:
    """What the work queue converges to when nodes are equally fast: client i → node i mod n,
    each node running its clients in order."""
    out: dict[int, list[int]] = {n: [] for n in node_ids}
    for i, cid in enumerate(sampled_clients):
        out[node_ids[i % len(node_ids)]].append(cid)
    return out


class ClientScheduler:
    """Dynamic work queue over nodes. ``dispatch(node_id, cid)`` sends one instruction,
    ``poll()`` returns finished ``(node_id, cid, reply)`` tuples (possibly empty)."""

    def __init__(self, sampled_clients: Iterable[int], node_ids: Iterable[int],
                 dispatch: Callable[[int, int], None], poll: Callable[[], list[tuple[int, int, Any]]],
                 poll_s: float = 0.0, is_alive: Callable[[int], bool] | None = None) -> None:
        """``is_alive(node_id)``: a node that is gone when its reply comes back (a remote node that stopped polling, a node whose
        workers died) leaves the rotation and the client it held goes to the next free node (ref: the node manager re-queues the
        client of a dead worker, photon/node_manager/node_manager_app.py:553-579; Flower drops silent nodes from ``get_node_ids``)."""
        self.queue = deque(sampled_clients)
        self.free = deque(node_ids)
        self.dispatch, self.poll, self.poll_s, self.is_alive = dispatch, poll, poll_s, is_alive
        self.in_flight: dict[int, int] = {}
        self.lost: list[int] = []

    def __iter__(self) -> Iterator[tuple[int, int, Any]]:
        while self.queue or self.in_flight:
            while self.queue and self.free:
                node, cid = self.free.popleft(), self.queue.popleft()
                self.in_flight[node] = cid
                self.dispatch(node, cid)
            for node, cid, reply in self.poll():
                held = self.in_flight.pop(node, None)
                if self.is_alive is not None and not self.is_alive(node):
                    self.lost.append(node)
                    if held is not None:
                        self.queue.appendleft(held)
                    continue
                self.free.append(node)
                yield node, cid, reply
            if self.queue and not self.free and not self.in_flight:
                raise RuntimeError(f"every node of the fleet is gone ({len(self.lost)} lost); {len(self.queue)} sampled client(s) were not trained")
            if self.poll_s:
                time.sleep(self.poll_s)


def fit_or_evaluate_ins(kind: str, server_round: int, client_ids: list[int], client_states: dict[int, ClientState],
                        server_steps_cumulative: int, per_client_config: dict[int, Any]) -> Message:
    """Instruction message carrying the reference's fields (ref: server_util.py:205-302)."""
    config = {"server_round": int(server_round), "client_ids": str(list(client_ids)),
              "client_state": encode_client_states(client_states), "server_steps_cumulative": int(server_steps_cumulative)}
    content: Any = FitIns(parameters=None, config=config) if kind == "train" else EvaluateIns(parameters=None, config=config)
    msg = Message(kind=kind, content=content, group_id=str(server_round))
    msg.per_client = {int(c): v for c, v in per_client_config.items()}  # type: ignore[attr-defined]
    return msg


fit_or_evaluate_ins_recordset = fit_or_evaluate_ins  # reference name (ref: server_util.py:205-302); no RecordSet here


# ----------------------------------------------------------------------------- reference names (ref: server_util.py:31,65-202)
from photon_b200.server.fit_utils import TooManyFailuresError  # noqa: E402,F401 - raised by the round loop; lives with the fit code here


def message_collaborative(send: Callable[[int, Message], None], receive: Callable[[], list[Message]], message_type: str,
                          sampled_clients: list[int], gen_ins_function: Callable[[int, int], Any], all_node_ids: list[int], current_round: int,
                          client_state: dict[int, ClientState], server_steps_cumulative: int, poll_s: float = 0.01) -> Iterator[Message]:
    """The reference's generator form of the work queue (ref: server_util.py:65-202): the first ``len(all_node_ids)`` sampled clients
    go out at once, every reply that comes back is yielded and frees its node for the next sampled client. ``send(node_id, message)``
    / ``receive() -> [replies]`` stand for the driver (``push_messages`` / ``pull_messages``); ``gen_ins_function(round, cid)`` builds
    the client's instruction config. :class:`ClientScheduler` is the loop underneath."""
    def dispatch(node: int, cid: int) -> None:
        msg = fit_or_evaluate_ins(message_type, current_round, [cid], client_state, server_steps_cumulative, {cid: gen_ins_function(current_round, cid)})
        msg.node_id = node
        send(node, msg)

    def poll() -> list[tuple[int, int, Message]]:
        return [(m.node_id, -1, m) for m in receive()]

    for _node, _cid, reply in ClientScheduler(sampled_clients, all_node_ids, dispatch, poll, poll_s=poll_s):
        yield reply

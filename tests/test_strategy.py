"""Server optimizers vs a NumPy oracle of SURVEY §2.4, streaming mean vs naive mean, metrics."""
import math

import numpy as np
import pytest
import torch

from photon_b200.config import compose
from photon_b200.messages import Code, EvaluateRes, FitRes, ParamHandle, Status
from photon_b200.strategy import (FedAdam, FedAvgEfficient, FedMom, FedNesterov, FedYogi, StreamingMean, dispatch_strategy,
                                  weighted_average)
from photon_b200.strategy.aggregation import naive_weighted_mean
from photon_b200.strategy.metrics import FedSimpleNoiseScale
from photon_b200.utils.flat import FlatLayout


def oracle(kind, x, a, m, v, hp, t, compat=False):
    x, a = x.astype(np.float64), a.astype(np.float64)
    pg = x - a
    if kind == "fedavg":
        return x - hp["lr"] * pg, m, v
    if kind == "nesterov":
        m = hp["mu"] * m + pg
        return x - hp["lr"] * (pg + hp["mu"] * m), m, v
    if kind == "fedmom":
        vn = x - hp["lr"] * pg
        return (1 + hp["mu"]) * vn - hp["mu"] * m, vn, v
    m = hp["beta1"] * m + (1 - hp["beta1"]) * pg
    if kind == "fedadam":
        v = hp["beta2"] * v + (1 - hp["beta2"]) * pg ** 2
    else:
        v = v + (1 - hp["beta2"]) * pg ** 2 * np.sign(pg ** 2 - v)
    step = hp["eta"] * (m / (1 - hp["beta1"] ** t)) / (np.sqrt(v / (1 - hp["beta2"] ** t)) + hp["tau"])
    return (x + step) if compat else (x - step), m, v


@pytest.mark.parametrize("cls,kw", [(FedAvgEfficient, dict(server_learning_rate=0.7)), (FedNesterov, dict(server_learning_rate=0.7, server_momentum=0.9)),
                                    (FedMom, dict(server_learning_rate=0.5, server_momentum=0.8)), (FedAdam, {}), (FedYogi, {})])
@pytest.mark.parametrize("compat", [False, True])
def test_server_optimizers_match_oracle(cls, kw, compat):
    rng = np.random.default_rng(0)
    n = 257
    s = cls(n_clients_per_round=3, reference_sign_compat=compat, **kw)
    x0 = rng.standard_normal(n).astype(np.float32)
    s.initialize(torch.from_numpy(x0.copy()))
    x, m, v = x0.astype(np.float64), np.zeros(n), np.zeros(n)
    for t in range(1, 4):
        clients = [x.astype(np.float32) + 0.1 * rng.standard_normal(n).astype(np.float32) for _ in range(3)]
        w = [10, 20, 30]
        a = sum(c.astype(np.float64) * wi for c, wi in zip(clients, w)) / sum(w)
        x, m, v = oracle(s.kind, x, a, m, v, s.hp, t, compat)
        res = [FitRes(Status(), ParamHandle("inline", torch.from_numpy(c)), wi, {"loss": 1.0}, i) for i, (c, wi) in enumerate(zip(clients, w))]
        p, metrics = s.aggregate_fit(t, iter(res))
        np.testing.assert_allclose(p.numpy(), x, rtol=2e-4, atol=2e-5)
        assert metrics["server/n_aggregated_clients"] == 3 and "server/l2_norm_pseudo_gradient" in metrics
    assert len(s.state_keys) == 1 + s.n_moments


def test_sign_default_descends():
    s = FedAdam(eta=0.1)
    x = torch.zeros(8)
    s.initialize(x.clone())
    target = torch.ones(8)
    for t in range(1, 30):
        s.aggregate_fit(t, iter([FitRes(Status(), ParamHandle("inline", s.parameters + 0.1 * (target - s.parameters)), 1, {}, 0)]))
    assert (s.parameters - target).abs().max() < (x - target).abs().max()   # moves TOWARDS the clients


def test_streaming_mean_equals_naive_and_gap_metric():
    torch.manual_seed(0)
    xs = [(torch.randn(1000), float(w)) for w in (3, 11, 7, 1)]
    sm = StreamingMean()
    for x, w in xs:
        sm.add(x, w)
    assert torch.allclose(sm.result(), naive_weighted_mean(xs), atol=1e-6)
    s = FedAvgEfficient(1.0, n_clients_per_round=4, track_inplace_aggregation=True)
    s.initialize(torch.zeros(1000))
    _, m = s.aggregate_fit(1, iter([FitRes(Status(), ParamHandle("inline", x), int(w), {}, i) for i, (x, w) in enumerate(xs)]))
    assert m["server/l2_norm_fedavg_gap"] < 1e-4


def test_scaling_fn_and_layerwise_norms():
    lay = FlatLayout.build([("a", (4, 4)), ("b", (8,))], align=4, total_multiple=4)
    s = FedNesterov(1.0, 0.0, n_clients_per_round=4, scaling_fn="sqrt")
    s.initialize(torch.zeros(lay.total), layout=lay)
    _, m = s.aggregate_fit(1, iter([FitRes(Status(), ParamHandle("inline", torch.ones(lay.total)), 1, {}, 0)]))
    assert math.isclose(s.scaling_factor(), 2.0)
    assert math.isclose(m["server/layer/0/l2_norm_fedavg_result"], 2.0 * 4.0, rel_tol=1e-6)
    with pytest.raises(ValueError):
        FedAvgEfficient(1.0, scaling_fn="cubic")


def test_weighted_average_merges_client_state():
    out = weighted_average([(10, {"loss": 1.0, "client_state_acc": str({0: {"steps_done": 2}})}),
                            (30, {"loss": 3.0, "client_state_acc": str({1: {"steps_done": 4}})})])
    assert math.isclose(out["loss"], 2.5) and "0:" in out["client_state_acc"] and "1:" in out["client_state_acc"]


def test_dispatcher_and_noise_scale():
    for name, cls in [("NESTOROV", FedNesterov), ("nesterov", FedNesterov), ("fedavg", FedAvgEfficient), ("fedmom", FedMom)]:
        assert isinstance(dispatch_strategy(compose([f"fl.strategy_name={name}"])), cls)
    s = dispatch_strategy(compose(["fl.strategy_name=fedyogi", "fl.strategy_kwargs={eta: 0.01, beta_1: 0.9, beta_2: 0.99, tau: 0.001}"]))
    assert isinstance(s, FedYogi) and s.hp["tau"] == 0.001
    assert dispatch_strategy(compose(["fl.strategy_name=fedavg"])).hp["lr"] == 1.0     # the reference pins eta = 1
    s = dispatch_strategy(compose(["fl.use_noise_scale_metric=true"]))
    assert isinstance(s.metrics_callback, FedSimpleNoiseScale)
    s.initialize(torch.zeros(64))
    res = [FitRes(Status(), ParamHandle("inline", torch.randn(64) * 0.1 + 1.0), 1, {}, i) for i in range(8)]
    _, m = s.aggregate_fit(1, iter(res))
    assert m["noise_scale/b_big"] == 8 and m["noise_scale/trace_estimate"] > 0


def test_aggregate_evaluate():
    s = FedAvgEfficient(1.0)
    loss, m = s.aggregate_evaluate(1, [EvaluateRes(Status(), 2.0, 10, {"ValLanguageCrossEntropy": 2.0}), EvaluateRes(Status(), 4.0, 30, {"ValLanguageCrossEntropy": 4.0}),
                                       EvaluateRes(Status(Code.FAILED), 0.0, 0, {})])
    assert math.isclose(loss, 3.5) and math.isclose(m["ValLanguageCrossEntropy"], 3.5)

"""Fused NVLink kernels on 1..8 GPUs (single process, peer access): federated round (all five server
optimizers) and the DDP all-reduce vs PyTorch references."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs >= {n} GPUs")
    return n


def _need2():
    return _need(2) and min(torch.cuda.device_count(), 4)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
@pytest.mark.parametrize("kind", ["fedavg", "nesterov", "fedmom", "fedadam", "fedyogi"])
def test_fed_round_matches_oracle(kind, n):
    """The fused round kernel on n GPUs of one process (n = 1: plain loads, bf16 cast from registers, no second pass)."""
    _need(n)
    from photon_b200.parallel.fed_round import NvlFedRound
    from photon_b200.strategy.strategies import FedAdam, FedAvgEfficient, FedMom, FedNesterov, FedYogi, server_opt_step
    from photon_b200.utils.flat import FlatLayout

    mk = {"fedavg": lambda: FedAvgEfficient(0.7), "nesterov": lambda: FedNesterov(0.7, 0.9), "fedmom": lambda: FedMom(0.5, 0.8),
          "fedadam": lambda: FedAdam(tau=1e-3), "fedyogi": lambda: FedYogi()}[kind]
    strat, ref = mk(), mk()
    # an uneven tensor table (tensors smaller and larger than a shard, one crossing every shard boundary)
    lay = FlatLayout.build([("a", (300, 1000)), ("b", (7,)), ("c", (513, 512)), ("d", (1 << 18,)), ("e", (1000, 200))])
    total = lay.total
    torch.manual_seed(0)
    x0 = torch.zeros(total)
    for i in range(len(lay.names)):
        lay.view(x0, i).normal_()
    fed = NvlFedRound(total, strat, devices=list(range(n)), layout=lay)
    fed.set_global(x0)
    ref.initialize(x0.clone(), layout=lay)
    for rnd in range(1, 4):
        fed.begin_round()
        clients, weights = [], []
        for g in range(n):
            for c in range(2):  # two multiplexed clients per GPU
                p = ref.parameters.clone()
                for i in range(len(lay.names)):
                    lay.view(p, i).add_(0.05 * torch.randn(lay.shapes[i]))
                w = float(10 * (g + 1) + c)
                fed.add_client(p.to(f"cuda:{g}"), w, local=g)
                clients.append(p), weights.append(w)
        fed.finish_round(rnd)
        for g in range(n):
            torch.cuda.synchronize(g)
        assert fed.check_status() == 0
        avg = (sum(p.double() * w for p, w in zip(clients, weights)) / sum(weights)).float()
        x_before = ref.parameters.clone()
        pg = server_opt_step(ref.kind, ref.parameters, avg, ref.momentum_vector, ref.second_momentum_vector, ref.hp, rnd)
        # Adam-type steps are eta * m_hat / (sqrt(v_hat) + tau): where |pg| is within a few tau of zero a 1-ulp difference of the fp32
        # mean moves the step by up to eta/tau ulps. Those elements are compared against that bound, everything else tightly.
        tight = torch.ones(total, dtype=torch.bool) if kind not in ("fedadam", "fedyogi") else (x_before - avg).abs() > 20 * ref.hp["tau"]
        assert float(tight.float().mean()) > 0.1
        for g in range(n):  # every GPU must hold the same new global model (fp32 + bf16 cast)
            got = fed.global_params(g).cpu()
            torch.testing.assert_close(got[tight], ref.parameters[tight], rtol=2e-5, atol=2e-6)
            # near pg = 0: |d step / d pg| <= eta / tau, and the fp32 weighted mean carries ~1e-7 * (1 + |x|) of rounding noise
            bound = ref.hp.get("eta", 0.0) / ref.hp.get("tau", 1.0) * 1e-6 * (1.0 + ref.parameters[~tight].abs()) + 1e-6
            if kind == "fedyogi":   # v += (1-b2) pg^2 sign(pg^2 - v) is discontinuous: a 1-ulp pg flips the sign where pg^2 ~ v -> v moves by 2 (1-b2) pg^2
                bound = bound + 2.0 * ref.hp["eta"] * (1.0 - ref.hp["beta2"])
            assert bool(((got[~tight] - ref.parameters[~tight]).abs() <= bound).all())
            assert torch.equal(fed.global_shadow(g).cpu(), got.to(torch.bfloat16))
        # server moments (each GPU keeps its shard): stitched they equal the oracle's planes
        for j, plane in enumerate((ref.momentum_vector, ref.second_momentum_vector)):
            if plane is None:
                continue
            full = sum(fed.full_moments(g)[j].cpu() for g in range(n))
            torch.testing.assert_close(full, plane, rtol=1e-4, atol=1e-6)
        # norm by-products of the same pass: global AND per tensor (the reference's server/layer/{i}/... metrics)
        want = ref.norm_metrics(pg, avg)
        # every round is compared as ONE step from identical state: adopt the kernel's state so that the (bounded) differences
        # near pg = 0 do not compound through the adaptive optimizers' history
        ref.parameters.copy_(fed.global_params(0).cpu())
        for j, name in enumerate(("momentum_vector", "second_momentum_vector")):
            if getattr(ref, name) is not None:
                getattr(ref, name).copy_(sum(fed.full_moments(g)[j].cpu() for g in range(n)))
        norms = fed.round_norms()
        assert set(want) == set(norms), sorted(set(want) ^ set(norms))[:5]
        for k, v in want.items():
            assert math.isclose(norms[k], v, rel_tol=1e-4, abs_tol=1e-6), (k, norms[k], v)
    fed.close()


def test_fed_round_failed_client_and_empty_round():
    n = _need2()
    from photon_b200.parallel.fed_round import NvlFedRound
    from photon_b200.strategy.strategies import FedAvgEfficient

    total = 1 << 16
    fed = NvlFedRound(total, FedAvgEfficient(1.0), devices=list(range(n)))   # no layout: the whole plane is one segment
    x0 = torch.randn(total)
    fed.set_global(x0)
    fed.begin_round()
    fed.add_client(torch.full((total,), 2.0, device="cuda:1"), 5.0, local=1)   # GPU 0's client "failed": contributes nothing
    fed.finish_round(1)
    for g in range(n):
        torch.cuda.synchronize(g)
    assert torch.allclose(fed.global_params(0).cpu(), torch.full((total,), 2.0))
    fed.begin_round()                                                           # nobody reports: model must be kept
    fed.finish_round(2)
    for g in range(n):
        torch.cuda.synchronize(g)
    assert torch.allclose(fed.global_params(1).cpu(), torch.full((total,), 2.0))
    fed.close()


def test_ddp_allreduce_kernel_multi_gpu():
    n = _need2()
    from photon_b200 import ops
    from photon_b200.parallel.symm import SymmArena

    total = (1 << 22) + 4096
    ar = SymmArena({"grads": (total, torch.float32)}, devices=list(range(n)))
    gs = [torch.randn(total, device=f"cuda:{g}") for g in range(n)]
    for g in range(n):
        ar.plane("grads", g).copy_(gs[g])
    norms = [torch.zeros(1, device=f"cuda:{g}") for g in range(n)]
    for it in range(2):
        ep = ar.next_epoch()
        for g in range(n):
            lo, hi = ar.shard(total, g)
            ops.ext().ddp_allreduce(ar.ctl_ptrs(), g, g, ep, ar.ptrs("grads"), lo, hi, norms[g])
        for g in range(n):
            torch.cuda.synchronize(g)
        ref = sum(x.cpu().double() for x in gs) / n if it == 0 else ref
        for g in range(n):
            assert torch.allclose(ar.plane("grads", g).cpu().double(), ref, rtol=1e-5, atol=1e-6)
            assert math.isclose(float(norms[g]), float(ref.norm()), rel_tol=1e-4)
    ar.close()

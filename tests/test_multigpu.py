"""Fused NVLink kernels across >= 2 GPUs (single process, peer access): federated round (all five server
optimizers) and the DDP all-reduce vs PyTorch references."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    return min(torch.cuda.device_count(), 4)


@pytest.mark.parametrize("kind", ["fedavg", "nesterov", "fedmom", "fedadam", "fedyogi"])
def test_fed_round_multi_gpu_matches_oracle(kind):
    n = _need2()
    from photon_b200.parallel.fed_round import NvlFedRound
    from photon_b200.strategy.strategies import FedAdam, FedAvgEfficient, FedMom, FedNesterov, FedYogi, server_opt_step

    mk = {"fedavg": lambda: FedAvgEfficient(0.7), "nesterov": lambda: FedNesterov(0.7, 0.9), "fedmom": lambda: FedMom(0.5, 0.8),
          "fedadam": lambda: FedAdam(tau=5e-2), "fedyogi": lambda: FedYogi()}[kind]
    # tau: Adam-type steps are eta*pg/(|pg|+tau) in round 1 - a 1-ulp difference in the fp32 mean is amplified by eta/tau near pg = 0
    strat, ref = mk(), mk()
    total = 1 << 20
    torch.manual_seed(0)
    x0 = torch.randn(total)
    fed = NvlFedRound(total, strat, devices=list(range(n)))
    fed.set_global(x0)
    ref.initialize(x0.clone())
    for rnd in range(1, 4):
        fed.begin_round()
        clients, weights = [], []
        for g in range(n):
            for c in range(2):  # two multiplexed clients per GPU
                p = (ref.parameters + 0.05 * torch.randn(total)).contiguous()
                w = float(10 * (g + 1) + c)
                fed.add_client(p.to(f"cuda:{g}"), w, local=g)
                clients.append(p), weights.append(w)
        fed.finish_round(rnd)
        for g in range(n):
            torch.cuda.synchronize(g)
        avg = sum(p.double() * w for p, w in zip(clients, weights)) / sum(weights)
        server_opt_step(ref.kind, ref.parameters, avg.float(), ref.momentum_vector, ref.second_momentum_vector, ref.hp, rnd)
        for g in range(n):  # every GPU must hold the same new global model (fp32 + bf16 cast)
            got = fed.global_params(g).cpu()
            bad = ((got - ref.parameters).abs() > 2e-5 + 2e-4 * ref.parameters.abs()).float().mean().item()
            assert bad < 1e-4, (kind, rnd, g, bad, (got - ref.parameters).abs().max())  # adaptive steps amplify 1-ulp pg differences
            assert torch.equal(fed.global_shadow(g).cpu(), got.to(torch.bfloat16))
        norms = fed.round_norms()
        assert math.isclose(norms["server/l2_norm_model"], float(ref.parameters.double().norm()), rel_tol=1e-3)
    fed.close()


def test_fed_round_failed_client_and_empty_round():
    n = _need2()
    from photon_b200.parallel.fed_round import NvlFedRound
    from photon_b200.strategy.strategies import FedAvgEfficient

    total = 1 << 16
    fed = NvlFedRound(total, FedAvgEfficient(1.0), devices=list(range(n)))
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

"""One-GPU federation through the fused round kernel (``photon.comm_stack.nvl`` with a world of one): virtual clients
multiplexed on the GPU, kernel by-product norms and the noise-scale estimate in the history, ``fl.aggregate_momenta``
(three planes through the same kernel), and agreement with the host transport."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = ["llm_config.model.d_model=256", "llm_config.model.n_heads=4", "llm_config.model.n_layers=2", "llm_config.max_seq_len=256",
        "llm_config.model.vocab_size=50368", "llm_config.global_train_batch_size=8", "llm_config.device_train_microbatch_size=4",
        "llm_config.local_steps=2ba", "llm_config.log_to_console=false", "~llm_config.loggers.wandb", "~llm_config.loggers.tensorboard",
        "~llm_config.callbacks", "llm_config.save_folder=null", "llm_config.optimizer.lr=1.0e-3", "llm_config.scheduler.schedulers.lr.t_warmup=1ba",
        "fl.eval_period=null", "photon.resume_round=null", "fl.n_total_clients=4", "fl.n_clients_per_round=3", "fl.n_rounds=2",
        "dataset.train.root_local=synthetic://c", "fl.use_noise_scale_metric=true", "fl.use_unigram_metrics=true"]


def _run(extra):
    from photon_b200.config import compose
    from photon_b200.federation import FederationRuntime
    from photon_b200.server_app import run_server

    cfg = compose(TINY + extra)
    rt = FederationRuntime(cfg, device=torch.device("cuda", 0), rank=0, world_size=1)
    h = run_server(cfg, runtime=rt)
    x = rt.round_backend.global_params().clone()
    total = rt.model_layout.total
    rt.close()
    return h.metrics_distributed_fit, x, total


def test_one_gpu_nvl_round_reports_norms_and_matches_host_transport():
    nvl, x_nvl, _ = _run(["run_uuid=g-nvl", "photon.comm_stack.shm=false", "photon.comm_stack.nvl=true"])
    ray, x_ray, _ = _run(["run_uuid=g-ray", "photon.comm_stack.shm=false", "photon.comm_stack.ray=true"])
    for fit in (nvl, ray):
        assert [v for _, v in fit["server/n_failures"]] == [0, 0]
        assert [v for _, v in fit["noise_scale/b_big"]] == [3, 3]
        assert any("Unigram" in k for k in fit)            # train-side unigram metrics travel with the fit metrics
    for key, tol in (("server/l2_norm_pseudo_gradient", 2e-2), ("server/l2_norm_model", 2e-3), ("server/l2_norm_momentum_vector", 2e-2)):
        a, b = nvl[key][-1][1], ray[key][-1][1]     # kernel by-products vs host norms (two bf16 training runs apart)
        assert abs(a - b) <= tol * abs(b) + 1e-6, (key, a, b)
    rel = float((x_nvl - x_ray).norm() / x_ray.norm())
    assert rel < 1e-3, rel


def test_one_gpu_nvl_round_with_aggregated_momenta():
    fit, x, total = _run(["run_uuid=g-mom", "photon.comm_stack.shm=false", "photon.comm_stack.nvl=true", "fl.aggregate_momenta=true",
                          "fl.reset_optimizer=false", "llm_config.optimizer.name=decoupled_adamw"])
    assert [v for _, v in fit["server/n_failures"]] == [0, 0]
    assert x.numel() == 3 * total and torch.isfinite(x).all()
    assert float(x[total:2 * total].abs().sum()) > 0 and float(x[2 * total:].min()) >= 0 and float(x[2 * total:].sum()) > 0

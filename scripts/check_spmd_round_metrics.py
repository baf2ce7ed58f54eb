"""torchrun --nproc-per-node 2 scripts/check_spmd_round_metrics.py — one nvl round on 2 GPUs; checks that the kernel by-product
norms and the noise-scale estimate (two NCCL all-reduces of a handful of scalars) reach rank 0's history and that every rank
ends with the same model. A quick multi-process probe of the metric collection path; the full comparison against the
NCCL transport lives in tests/test_multiproc_gpu.py."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from photon_b200.config import compose
    from photon_b200.federation import FederationRuntime
    from photon_b200.server_app import run_server

    tiny = ["llm_config.model.d_model=256", "llm_config.model.n_heads=4", "llm_config.model.n_layers=1", "llm_config.max_seq_len=256",
            "llm_config.global_train_batch_size=4", "llm_config.device_train_microbatch_size=4", "llm_config.local_steps=3ba",
            "llm_config.optimizer.lr=1.0e-3", "llm_config.scheduler.schedulers.lr.t_warmup=1ba",   # ADOPT only seeds v on its first step
            "llm_config.log_to_console=false", "~llm_config.loggers.wandb", "~llm_config.loggers.tensorboard", "~llm_config.callbacks",
            "llm_config.save_folder=null", "fl.eval_period=null", "photon.resume_round=null", "fl.n_total_clients=4", "fl.n_clients_per_round=4",
            "fl.n_rounds=1", "dataset.train.root_local=synthetic://c", "fl.use_noise_scale_metric=true", "photon.comm_stack.shm=false",
            "photon.comm_stack.nvl=true", "run_uuid=probe"]
    t0 = time.time()
    cfg = compose(tiny + sys.argv[1:])
    rt = FederationRuntime(cfg, device=torch.device("cuda", lr), rank=dist.get_rank(), world_size=dist.get_world_size())
    h = run_server(cfg, runtime=rt)
    x = rt.round_backend.global_params().clone()
    ref = x.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, x), "ranks hold different global models"
    if dist.get_rank() == 0:
        fit = h.metrics_distributed_fit
        keys = ["server/n_failures", "server/l2_norm_pseudo_gradient", "server/l2_norm_model", "noise_scale/b_big", "noise_scale/noise_scale_raw"]
        vals = {k: fit[k][-1][1] for k in keys}
        assert vals["server/n_failures"] == 0 and vals["noise_scale/b_big"] == 4 and vals["server/l2_norm_pseudo_gradient"] > 0, vals
        print("PROBE_OK", {k: round(float(v), 5) for k, v in vals.items()}, f"{time.time() - t0:.1f}s", flush=True)
    rt.close()
    dist.barrier()
    dist.destroy_process_group()

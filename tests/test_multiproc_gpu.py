"""torchrun (one process per GPU) paths: centralised DDP with the fused NVLink all-reduce (config #3), the SPMD
federation with the nvl transport, and 2 GPUs per client (config #5). Need >= 2 GPUs."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest
import torch
from conftest import free_port as _free_port  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
TINY = ["llm_config.model.d_model=256", "llm_config.model.n_heads=4", "llm_config.model.n_layers=2", "llm_config.max_seq_len=256",
        "llm_config.model.vocab_size=50368", "llm_config.global_train_batch_size=8", "llm_config.device_train_microbatch_size=4",
        "llm_config.local_steps=2ba", "llm_config.log_to_console=false", "~llm_config.loggers.wandb", "~llm_config.loggers.tensorboard",
        "~llm_config.callbacks", "llm_config.save_folder=null", "llm_config.optimizer.lr=1.0e-3", "llm_config.scheduler.schedulers.lr.t_warmup=1ba",
        "fl.eval_period=null", "photon.resume_round=null"]


def _torchrun(tmp_path, n, body, port, env=None):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs >= {n} GPUs")
    script = tmp_path / "run.py"
    script.write_text("import os, sys, torch, torch.distributed as dist\n" f"sys.path.insert(0, {str(ROOT)!r})\n"
                      "torch.cuda.set_device(int(os.environ['LOCAL_RANK']))\n"
                      "dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))\n"
                      f"TINY = {TINY!r}\n" + textwrap.dedent(body) + "\ndist.barrier(); dist.destroy_process_group()\n")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                          "--master-port", str(_free_port() if port else port), str(script)], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1", **(env or {})))
    assert out.returncode == 0 and "RESULT_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-4000:]


def test_centralised_ddp_fused_allreduce_matches_nccl(tmp_path):
    """DDP inside one job: fused NVLink all-reduce (replicated optimizer), fused ZeRO step (fsdp_config: reduce-scatter +
    clip + optimizer shard + all-gather in one kernel) and NCCL + sharded-state broadcast all end on the same parameters."""
    _torchrun(tmp_path, 2, """
        from photon_b200.config import compose
        from photon_b200.centralised_train import run_centralised
        rank = dist.get_rank(); dev = torch.device('cuda', rank)
        outs = {}
        for name, nvl, extra, cls in (('zero', True, [], 'NvlZeroComm'), ('nvl', True, ['~llm_config.fsdp_config'], 'NvlGradComm'),
                                      ('nccl', False, ['~llm_config.fsdp_config'], 'NcclGradComm'),
                                      ('nccl_sharded', False, [], 'NcclGradComm')):
            cfg = compose(TINY + ['run_uuid=ddp', 'dataset/streams@dataset.train.streams=centralised', 'dataset.train.root_local=synthetic://7'] + extra)
            tr = run_centralised(cfg, device=dev, rank=rank, world_size=2, duration='3ba', use_nvl_allreduce=nvl)
            assert type(tr.grad_comm).__name__ == cls, (name, type(tr.grad_comm).__name__)
            assert tr.state.optimizer.sharded == (name in ('zero', 'nccl_sharded')), name
            x = tr.state.flat.params.clone()
            ref = x.clone(); dist.broadcast(ref, src=0)
            assert torch.equal(ref, x), name + ': ranks diverged'
            if name == 'zero':   # the bf16 compute copy written by the peers equals the cast of the masters
                assert torch.equal(tr.state.backend.bf16_params, x.to(torch.bfloat16))
            outs[name] = x; tr.close()
        rels = {k: ((outs[k] - outs['nccl']).norm() / outs['nccl'].norm()).item() for k in outs}
        assert all(r < 2e-3 for r in rels.values()), rels
        if rank == 0: print('RESULT_OK', rels)
    """, 29541)


def test_spmd_federation_nvl_and_two_gpus_per_client(tmp_path):
    _torchrun(tmp_path, 2, """
        from photon_b200.config import compose
        from photon_b200.federation import FederationRuntime
        from photon_b200.server_app import run_server
        rank = dist.get_rank(); dev = torch.device('cuda', rank)
        finals = {}
        for name, extra, gpc in (('nvl', ['photon.comm_stack.shm=false', 'photon.comm_stack.nvl=true'], 1),
                                 ('ray', ['photon.comm_stack.shm=false', 'photon.comm_stack.ray=true'], 1),
                                 ('nvl2', ['photon.comm_stack.shm=false', 'photon.comm_stack.nvl=true'], 2)):
            cfg = compose(TINY + ['run_uuid=fed-' + name, 'fl.n_total_clients=4', 'fl.n_clients_per_round=4', 'fl.n_rounds=2',
                                  'fl.strategy_name=fedadam', 'fl.strategy_kwargs={eta: 0.01, beta_1: 0.9, beta_2: 0.99, tau: 0.001}',
                                  'fl.use_noise_scale_metric=true', 'dataset.train.root_local=synthetic://c'] + extra)
            rt = FederationRuntime(cfg, device=dev, rank=rank, world_size=2, gpus_per_client=gpc)
            h = run_server(cfg, runtime=rt)
            x = rt.round_backend.global_params().clone()
            ref = x.clone(); dist.broadcast(ref, src=0)
            assert torch.equal(ref, x), name + ': ranks hold different global models'
            finals[name] = x
            if rank == 0:
                assert h.latest('server/n_failures') == 0, (name, h.metrics_distributed_fit)
                fit = h.metrics_distributed_fit    # kernel by-product norms + the noise-scale estimate reach the history on every transport
                assert len(fit['server/l2_norm_pseudo_gradient']) == 2 and fit['noise_scale/b_big'][-1][1] == 4, name
                finals[name + '/pg'] = fit['server/l2_norm_pseudo_gradient'][-1][1]
            rt.close()
        rel = ((finals['nvl'] - finals['ray']).norm() / finals['ray'].norm()).item()
        assert rel < 1e-3, rel          # fused NVLink round == NCCL all-reduce baseline
        assert torch.isfinite(finals['nvl2']).all()
        if rank == 0:
            assert abs(finals['nvl/pg'] - finals['ray/pg']) < 2e-2 * finals['ray/pg'], (finals['nvl/pg'], finals['ray/pg'])
            print('RESULT_OK', rel)
    """, 29543)


def test_nvls_multimem_paths_match_p2p(tmp_path):
    """NVLS (multimem.ld_reduce / multimem.st on the multicast mapping of the arena) against the P2P loops of the same
    kernels: federated round (FedAdam) and the fused ZeRO step. Skips itself when the box has no multicast support."""
    _torchrun(tmp_path, 2, """
        import os
        from photon_b200.config import compose
        from photon_b200.federation import FederationRuntime
        from photon_b200.server_app import run_server
        from photon_b200.centralised_train import run_centralised
        rank = dist.get_rank(); dev = torch.device('cuda', rank)
        res = {}
        for mode in ('nvls', 'p2p'):
            os.environ['PB_NVLS'] = '1' if mode == 'nvls' else '0'
            cfg = compose(TINY + ['run_uuid=nv-' + mode, 'fl.n_total_clients=4', 'fl.n_clients_per_round=4', 'fl.n_rounds=2',
                                  'fl.strategy_name=fedadam', 'fl.strategy_kwargs={eta: 0.01, beta_1: 0.9, beta_2: 0.99, tau: 0.001}',
                                  'dataset.train.root_local=synthetic://c', 'photon.comm_stack.shm=false', 'photon.comm_stack.nvl=true'])
            rt = FederationRuntime(cfg, device=dev, rank=rank, world_size=2)
            run_server(cfg, runtime=rt)
            res[mode] = (rt.round_backend.global_params().clone(), bool(rt.round_backend.fed.arena.mc_ptr('acc')))
            rt.close()
            cfg = compose(TINY + ['run_uuid=nvz-' + mode, 'dataset/streams@dataset.train.streams=centralised', 'dataset.train.root_local=synthetic://7'])
            tr = run_centralised(cfg, device=dev, rank=rank, world_size=2, duration='3ba', use_nvl_allreduce=True)
            res['z' + mode] = (tr.state.flat.params.clone(), bool(tr.grad_comm.arena.mc_ptr('grads')))
            tr.close()
        if not res['nvls'][1]:
            if rank == 0: print('RESULT_OK (no multicast on this box: nothing to compare)')
        else:
            assert res['znvls'][1] and not res['p2p'][1] and not res['zp2p'][1]
            for a, b in (('nvls', 'p2p'), ('znvls', 'zp2p')):
                rel = ((res[a][0] - res[b][0]).norm() / res[b][0].norm()).item()
                # the switch adds the ranks in its own order: 1-ulp differences in the mean, amplified by the Adam-type steps
                assert rel < 2e-3, (a, rel)
            if rank == 0: print('RESULT_OK nvls == p2p')
    """, 29545, env={"PB_NVLS_MIN_WORLD": "2"})


@pytest.mark.parametrize("kind", ["kill"])
def test_rank_killed_mid_round_nvl_round_completes_over_survivors(tmp_path, kind):
    """2 GPUs, comm_stack.nvl: rank 1 is SIGKILLed while it trains in round 2. Rank 0 rules it out (heartbeat), runs the fused round
    kernel over the survivors (subset control pages, P2P path, the dead rank's server shard adopted), counts the lost clients as
    failures and finishes round 3 alone. Started with photon_b200.launch (torchrun would kill the survivor too)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "run.py"
    script.write_text("import os, sys, math, torch, torch.distributed as dist\n" f"sys.path.insert(0, {str(ROOT)!r})\n" f"TINY = {TINY!r}\n" + textwrap.dedent(f"""
        torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
        from photon_b200.config import compose
        from photon_b200.federation import FederationRuntime
        from photon_b200.server_app import run_server
        rank = dist.get_rank(); dev = torch.device('cuda', rank)
        cfg = compose(TINY + ['run_uuid=ftgpu', 'fl.n_total_clients=4', 'fl.n_clients_per_round=4', 'fl.n_rounds=3', 'fl.accept_failures_cnt=4',
                              'fl.strategy_name=fedadam', 'fl.strategy_kwargs={{eta: 0.01, beta_1: 0.9, beta_2: 0.99, tau: 0.001}}',
                              'dataset.train.root_local=synthetic://c', 'photon.comm_stack.shm=false', 'photon.comm_stack.nvl=true',
                              'photon.liveness_timeout_s=3', 'photon.kernel_peer_timeout_s=5',
                              'fl.fault_injection={{round: 2, rank: 1, kind: {kind}}}'])
        rt = FederationRuntime(cfg, device=dev, rank=rank, world_size=2)
        h = run_server(cfg, runtime=rt)
        if rank == 0:
            fit = h.metrics_distributed_fit
            fails = dict(fit['server/n_failures'])
            assert fails[1] == 0 and fails[2] >= 1 and fails[3] == 0, fails
            assert 'server/round_ignored' not in fit
            assert [r for r, _ in fit['server/l2_norm_pseudo_gradient']] == [1, 2, 3]
            assert dict(fit['server/n_nodes'])[3] == 1 and rt.alive_ranks == [0]
            x = rt.round_backend.global_params()
            assert bool(torch.isfinite(x).all()) and torch.equal(rt.round_backend.global_shadow(), x.to(torch.bfloat16))
            print('RESULT_OK survivors', rt.alive_ranks)
        os._exit(0)
    """))
    out = subprocess.run([sys.executable, "-m", "photon_b200.launch", "--nproc", "2", "--master-port", str(_free_port()), str(script)],
                         capture_output=True, text=True, timeout=900, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), cwd=str(ROOT))
    assert out.returncode == 0 and "RESULT_OK survivors [0]" in out.stdout, out.stdout[-3000:] + out.stderr[-4000:]

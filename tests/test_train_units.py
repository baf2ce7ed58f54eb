"""Timestamp / schedulers / optimizers (closed forms) / metrics / trainer checkpoints — CPU."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from photon_b200.metrics import build_metrics, unigram_log_probs, unigram_loss_sum
from photon_b200.models.mpt import MPTConfig, MPTForCausalLM, shift_labels, trainable_named_parameters
from photon_b200.train.optim import ADOPT, DecoupledAdamW, build_optimizer, clip_coefficient
from photon_b200.train.schedulers import build_scheduler
from photon_b200.train.timestamp import Time, Timestamp
from photon_b200.utils.flat import FlatLayout, FlatParams
from conftest import free_port as _free_port  # noqa: E402


def test_time_strings():
    assert Time.parse("500ba").to_batches() == 500 and str(Time.parse("1e3ba")) == "1000ba"
    assert Time.parse("1ep").to_batches(batches_per_epoch=7) == 7
    assert Time.parse("1024sp").to_batches(samples_per_batch=256) == 4
    assert Time.parse("0.5dur").to_batches(max_duration=Time.parse("4800ba")) == 2400
    assert Time.parse("1e6tok").to_batches(tokens_per_batch=65536) == 15
    with pytest.raises(ValueError):
        Time.parse("12 parsecs")
    ts = Timestamp()
    ts.advance_batch(32, 65536)
    sd = ts.state_dict()
    t2 = Timestamp()
    t2.load_state_dict(sd)
    assert t2 == ts and t2.batch == 1 and t2.token == 65536


def test_schedulers_closed_form():
    f = build_scheduler({"schedulers": {"lr": {"name": "cosine_with_warmup", "t_warmup": "100ba", "alpha_f": 0.1, "t_max": "4800ba"}}})
    assert f(0) == 0.0 and math.isclose(f(50), 0.5) and math.isclose(f(100), 1.0)
    assert math.isclose(f(2450), 0.1 + 0.9 * 0.5, rel_tol=1e-6) and math.isclose(f(4800), 0.1) and math.isclose(f(9999), 0.1)
    g = build_scheduler({"name": "constant_with_sqrt_cooldown_with_warmup", "t_warmup": "100ba", "t_max": "5120ba", "t_cooldown": "240ba"})
    assert math.isclose(g(100), 1.0) and math.isclose(g(4880), 1.0) and math.isclose(g(4880 + 60), 1 - math.sqrt(0.25)) and math.isclose(g(5120), 0.0, abs_tol=1e-9)
    h = build_scheduler({"name": "linear_decay_with_warmup", "t_warmup": "0ba", "t_max": "10ba", "alpha_f": 0.0})
    assert math.isclose(h(5), 0.5)


def _flat(n=64):
    torch.manual_seed(0)
    m = nn.Linear(n, 1, bias=False)
    return FlatParams(m)


def test_adopt_closed_form():
    fp = _flat()
    opt = ADOPT(fp, lr=0.1, betas=(0.9, 0.99), eps=1e-6, use_kernel=False)
    n = fp.layout.numels[0]
    p0 = fp.params.clone()
    g1, g2 = torch.full((fp.layout.total,), 2.0), torch.full((fp.layout.total,), -1.0)
    fp.grads.copy_(g1)
    opt.step()
    assert torch.equal(fp.params, p0) and torch.allclose(opt.exp_avg_sq[:n], torch.full((n,), 4.0))    # step 0 only seeds v
    fp.grads.copy_(g2)
    opt.step()
    ng = max(-1.0 / 2.0, -1.0)          # g / sqrt(v) = -0.5, clip = 1**0.25 = 1
    m = 0.1 * ng
    assert torch.allclose(fp.params[:n], p0[:n] - 0.1 * m, atol=1e-7)
    assert torch.allclose(opt.exp_avg_sq[:n], torch.full((n,), 0.99 * 4 + 0.01 * 1.0))


def test_decoupled_adamw_matches_torch_adamw_without_decay():
    fp = _flat()
    ref = torch.nn.Parameter(fp.params[: fp.layout.numels[0]].clone())
    topt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    opt = DecoupledAdamW(fp, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, use_kernel=False)
    for i in range(5):
        g = torch.randn(fp.layout.total)
        fp.grads.copy_(g)
        ref.grad = g[: ref.numel()].clone()
        opt.step()
        topt.step()
    assert torch.allclose(fp.params[: ref.numel()], ref.detach(), atol=1e-6)
    o2 = build_optimizer({"name": "decoupled_adamw", "lr": 1.0, "weight_decay": 0.1}, _flat(), use_kernel=False)
    p = o2.flat.params.clone()
    o2.flat.grads.zero_()
    o2.step(0.5)                                          # lr factor 0.5 -> decay 1 - 0.5*0.1
    assert torch.allclose(o2.flat.params, p * 0.95, atol=1e-6)
    assert math.isclose(float(clip_coefficient(torch.tensor(4.0), 1.0)), 1.0 / (4.0 + 1e-6), rel_tol=1e-6)


def test_flat_layout_sorted_and_roundtrip():
    cfg = MPTConfig(d_model=32, n_heads=2, n_layers=11, max_seq_len=16, vocab_size=64, attn_impl="torch")
    model = MPTForCausalLM(cfg, seed=1)
    names = [n for n, _ in trainable_named_parameters(model)]
    assert names == sorted(names) and names.index("transformer.blocks.10.attn.Wqkv.bias") < names.index("transformer.blocks.2.attn.Wqkv.bias")
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    fp = FlatParams(model)
    assert all(torch.equal(before[n], p) for n, p in model.named_parameters())       # re-pointing keeps values
    assert all(o % 256 == 0 for o in fp.layout.offsets) and fp.layout.n_params == cfg.num_params()
    arrays = fp.to_ndarrays()
    fp.params.zero_()
    fp.load_ndarrays(arrays)
    assert all(torch.equal(before[n], p) for n, p in model.named_parameters())
    assert MPTConfig().num_params() == 125_311_488
    with pytest.raises(ValueError):
        fp.load_ndarrays(arrays[:-1])


def test_mpt_loss_shift_and_tied_head():
    cfg = MPTConfig(d_model=32, n_heads=2, n_layers=1, max_seq_len=8, vocab_size=50, attn_impl="torch")
    model = MPTForCausalLM(cfg, seed=0)
    ids = torch.randint(0, 50, (2, 8))
    t = shift_labels(ids)
    assert torch.equal(t[:, :-1], ids[:, 1:]) and (t[:, -1] == -100).all()
    loss, n = model.loss(ids)
    assert int(n) == 2 * 7 and abs(float(loss) - math.log(50)) < 1.5
    logits = model(ids)
    assert logits.shape == (2, 8, 50)
    # causal: changing a future token must not change past logits
    ids2 = ids.clone()
    ids2[:, -1] = (ids2[:, -1] + 1) % 50
    assert torch.allclose(model(ids2)[:, :-1], logits[:, :-1], atol=1e-5)
    for variant in (dict(alibi=True), dict(rope=True)):
        m2 = MPTForCausalLM(MPTConfig(d_model=32, n_heads=2, n_layers=1, max_seq_len=8, vocab_size=50, attn_impl="torch", learned_pos_emb=False, **variant), seed=0)
        assert torch.isfinite(m2.loss(ids)[0])


def test_metrics_and_unigram():
    ms = build_metrics(True)
    lp = unigram_log_probs({"1": 90, "2": 10}, 4)
    tg = torch.tensor([1, 1, 2, -100])
    uni = float(unigram_loss_sum(tg, lp))
    for m in ms.values():
        m.update({"loss_sum": 6.0, "n_tokens": 3.0, "n_correct": 2.0, "unigram_loss_sum": uni})
    assert math.isclose(ms["LanguageCrossEntropy"].compute(), 2.0) and math.isclose(ms["LanguagePerplexity"].compute(), math.exp(2.0))
    assert math.isclose(ms["TokenAccuracy"].compute(), 2 / 3)
    assert math.isclose(ms["UnigramNormalizedLanguageCrossEntropy"].compute(), 2.0 - uni / 3)
    assert math.isclose(ms["PureUnigramPerplexity"].compute(), math.exp(uni / 3))


def test_trainer_checkpoint_resume_and_ignore_keys(tmp_path):
    from photon_b200.data.streaming import Stream, StreamingTokenDataset, TokenLoader
    from photon_b200.train.trainer import Trainer

    cfg = MPTConfig(d_model=32, n_heads=2, n_layers=1, max_seq_len=16, vocab_size=50277, attn_impl="torch")

    def mk():
        ds = StreamingTokenDataset([Stream(local="synthetic://3")], seq_len=16, synthetic_samples=64)
        return Trainer(cfg, optimizer_cfg=dict(name="adopt", lr=1e-2, betas=[0.9, 0.99], eps=1e-6), train_loader=TokenLoader(ds, 4),
                       scheduler_cfg=dict(name="cosine_with_warmup", t_warmup="2ba", t_max="20ba", alpha_f=0.1),
                       global_train_batch_size=4, device_train_microbatch_size=2, precision="fp32", max_duration="20ba", grad_clip_norm=1.0,
                       device="cpu", save_folder=str(tmp_path), save_interval="3ba", save_num_checkpoints_to_keep=1, seed=5)

    a = mk()
    a.fit("6ba")
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == ["ep0-ba6-rank0.pt", "latest-rank0.pt"]                 # keep=1 removed ba3
    a.fit("2ba")
    b = mk()
    b.load_checkpoint(tmp_path / "latest-rank0.pt")
    assert b.state.timestamp.batch == 6 and b.state.optimizer.step_count == 6
    b.fit("2ba")                                                            # same data position + state => same weights
    assert torch.allclose(a.state.flat.params, b.state.flat.params, atol=1e-6)
    c = mk()
    c.load_checkpoint(tmp_path / "latest-rank0.pt", ignore_keys=["*optim*", "*timestamp*"])
    assert c.state.optimizer.step_count == 0 and c.state.timestamp.batch == 0
    saved = torch.load(tmp_path / "latest-rank0.pt", weights_only=False)["state"]["model"]
    lay = c.state.flat.layout
    assert all(torch.equal(lay.view(c.state.flat.params, n), saved[n]) for n in lay.names)


def test_sharded_optimizer_state_matches_replicated(tmp_path):
    """fsdp_config → ZeRO-style state sharding: 2 gloo ranks, each keeps half of the moments, end on the same
    parameters as the replicated-state run (ref: conf/llm_config/mpt-125m.yaml:85-91)."""
    import subprocess
    import sys
    import textwrap
    from pathlib import Path

    script = tmp_path / "zero.py"
    script.write_text(textwrap.dedent(f"""
        import sys, torch, torch.distributed as dist
        sys.path.insert(0, {str(Path(__file__).resolve().parents[1])!r})
        from photon_b200.models.mpt import MPTConfig
        from photon_b200.train.trainer import Trainer, shard_bounds
        dist.init_process_group("gloo")
        r = dist.get_rank()
        cfg = MPTConfig(d_model=32, n_heads=2, n_layers=2, max_seq_len=16, vocab_size=64, attn_impl="torch")
        def make(shard):
            return Trainer(cfg, optimizer_cfg=dict(name="adopt", lr=1e-2, betas=(0.9, 0.99), eps=1e-6),
                           scheduler_cfg=dict(name="constant_with_warmup", t_warmup="0ba"), global_train_batch_size=4, device_train_microbatch_size=2,
                           precision="fp32", device="cpu", rank=r, world_size=2, seed=3, shard_optimizer_state=shard,
                           kernels=dict(gemm="torch", attention="torch", norm="torch", loss="torch", optimizer="torch"))
        a, b = make(False), make(True)
        assert b.state.optimizer.sharded and b.state.optimizer.exp_avg.numel() < a.state.optimizer.exp_avg.numel()
        bounds = shard_bounds(a.state.flat.params.numel(), 2)
        assert bounds[0][1] == bounds[1][0] and bounds[1][1] == a.state.flat.params.numel() and bounds[0][1] % 256 == 0
        g = torch.Generator().manual_seed(100 + r)
        for _ in range(4):
            batch = dict(input_ids=torch.randint(0, 64, (2, 16), generator=g))
            a._train_batch(batch); b._train_batch(batch)
        assert torch.allclose(a.state.flat.params, b.state.flat.params, atol=1e-6), (a.state.flat.params - b.state.flat.params).abs().max()
        m_full, v_full = b.state.optimizer.full_moments()
        assert torch.allclose(m_full, a.state.optimizer.exp_avg, atol=1e-6) and torch.allclose(v_full, a.state.optimizer.exp_avg_sq, atol=1e-7)
        sd = b.state.optimizer.state_dict(); b.state.optimizer.load_state_dict(sd)          # per-rank shard round trip
        b.state.optimizer.load_state_dict(a.state.optimizer.state_dict())                    # full-state ckpt -> sharded optimizer
        if r == 0: print("OK")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(_free_port()), str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_profiler_callback_exports_chrome_traces(tmp_path):
    """llm_config.profiler → cyclic schedule + JSON trace handler (ref: clients/trainer_utils.py:1456-1482)."""
    from photon_b200.models.mpt import MPTConfig
    from photon_b200.train.callbacks import build_callbacks
    from photon_b200.train.trainer import Trainer

    cfg = MPTConfig(d_model=32, n_heads=2, n_layers=1, max_seq_len=16, vocab_size=64, attn_impl="torch")
    ids = torch.randint(0, 64, (2, 16))

    class Loader:
        def __iter__(self):
            while True:
                yield {"input_ids": ids}

    cbs = build_callbacks({"profiler": {"schedule": {"skip_first": 1, "wait": 0, "warmup": 1, "active": 2, "repeat": 1},
                                        "json_trace_handler": {"folder": str(tmp_path / "traces")}}})
    tr = Trainer(cfg, optimizer_cfg=dict(name="sgd", lr=1e-2), scheduler_cfg=dict(name="constant_with_warmup", t_warmup="0ba"),
                 train_loader=Loader(), global_train_batch_size=2, device_train_microbatch_size=2, precision="fp32", device="cpu",
                 callbacks=cbs, kernels=dict(gemm="torch", attention="torch", norm="torch", loss="torch", optimizer="torch"))
    tr.fit("6ba")
    traces = list((tmp_path / "traces").glob("rank0.*.pt.trace.json"))
    assert len(traces) == 1 and traces[0].stat().st_size > 1000
    tr.close()


def test_fsdp_config_selects_the_sharded_step():
    """fsdp_config (YAML default) → fused ZeRO step / sharded optimizer; NO_SHARD or a deleted node → plain DDP."""
    from photon_b200.config import compose
    from photon_b200.parallel.ddp import wants_sharded_step

    cfg = compose(["run_uuid=t"])
    assert wants_sharded_step(cfg["llm_config"])                               # shipped YAML: FSDP FULL_SHARD
    assert not wants_sharded_step(compose(["run_uuid=t", "~llm_config.fsdp_config"])["llm_config"])
    assert not wants_sharded_step(compose(["run_uuid=t", "llm_config.fsdp_config.sharding_strategy=NO_SHARD"])["llm_config"])
    assert wants_sharded_step(compose(["run_uuid=t", "llm_config.fsdp_config.sharding_strategy=SHARD_GRAD_OP"])["llm_config"])


# ------------------------------------------------------------------ reference-named API surface
def _tiny_trainer(tmp_path):
    from photon_b200.train.trainer import Trainer

    cfg = MPTConfig(d_model=32, n_heads=2, n_layers=1, max_seq_len=16, vocab_size=64, attn_impl="torch")
    ids = torch.randint(0, 64, (2, 16))

    class Loader:
        def __iter__(self):
            while True:
                yield {"input_ids": ids}

    tr = Trainer(cfg, optimizer_cfg=dict(name="decoupled_adamw", lr=1e-3), scheduler_cfg=dict(name="constant_with_warmup", t_warmup="0ba"),
                 train_loader=Loader(), global_train_batch_size=2, device_train_microbatch_size=2, precision="fp32", device="cpu",
                 kernels=dict(gemm="torch", attention="torch", norm="torch", loss="torch", optimizer="torch"))
    tr.fit("2ba")
    return tr



def test_param_side_channel_roundtrip(tmp_path):
    """Sender/receiver halves of the parameter side channel over shm and npz files."""
    import uuid

    from photon_b200.messages import ParamHandle
    from photon_b200.server.s3_utils import (release_remote_parameters, replace_parameters_in_recordset_with_remote,
                                             replace_remote_with_parameters_in_recordset)
    from photon_b200.utils.flat import FlatLayout

    lay = FlatLayout.build([("b.weight", (3, 5)), ("a.bias", (7,)), ("c", ())], align=8, total_multiple=16)
    flat = torch.arange(lay.total, dtype=torch.float32)
    want = lay.to_ndarrays(flat)
    for stack in ({"shm": True}, {"s3": True}):
        h = replace_remote_with_parameters_in_recordset(ParamHandle("inline", flat), stack, endpoint_id=f"pbt{uuid.uuid4().hex[:8]}",
                                                        layout=lay, root=tmp_path)
        assert h.kind == ("shm" if "shm" in stack else "file") and not torch.is_tensor(h.data)
        got = replace_parameters_in_recordset_with_remote(h)
        assert all(np.array_equal(a, b) for a, b in zip(got.data, want))
        back = replace_parameters_in_recordset_with_remote(h, layout=lay, as_flat=True).data
        assert all(np.array_equal(a, b) for a, b in zip(lay.to_ndarrays(back), want))
        release_remote_parameters(h)
    assert replace_remote_with_parameters_in_recordset(ParamHandle("inline", flat), {"nvl": True}, endpoint_id="x").kind == "nvl"


def test_reference_named_helpers(tmp_path):
    from photon.metrics.unigram_normalized_metrics import create_wrapped_subclass
    from photon_b200.strategy import (FedAdam, FedAvgEfficient, aggregate_parameters, initialize_strategy,
                                      parameters_to_ndarrays_gen)
    from photon_b200.utils.core import (chunks_idx, get_parameters_from_state, get_trainable_params_dict,
                                        get_wte_parameters_from_trainer, is_literal_for_ast, l2_norm_of_momenta,
                                        set_trainer_params_from_ndarrays, set_wte_parameters_to_trainer)
    from photon_b200.utils.flat import FlatLayout

    # running mean, functional form == textbook weighted mean
    xs, ws = [torch.randn(64) for _ in range(3)], [3.0, 1.0, 4.0]
    acc, n = None, 0.0
    for x, w in zip(xs, ws):
        acc, n = aggregate_parameters(acc, x, n, w)
    assert torch.allclose(acc, sum(x * w for x, w in zip(xs, ws)) / sum(ws), atol=1e-6) and n == 8.0
    lay = FlatLayout.build([("w", (4, 4)), ("b", (4,))], align=8, total_multiple=16)
    flat = torch.randn(lay.total)
    assert [a.shape for a in parameters_to_ndarrays_gen(flat, lay)] == [(4,), (4, 4)]
    # strategy initialisation: momenta required exactly when the strategy uses them
    initialize_strategy(FedAvgEfficient(1.0), flat.clone(), layout=lay)
    with pytest.raises(ValueError):
        initialize_strategy(FedAdam(), flat.clone(), layout=lay)
    st = FedAdam()
    initialize_strategy(st, lay.to_ndarrays(flat), lay.to_ndarrays(torch.zeros(lay.total)), lay.to_ndarrays(torch.zeros(lay.total)), layout=lay)
    assert torch.equal(lay.view(st.parameters, "w"), lay.view(flat, "w")) and st.second_momentum_vector is not None
    # misc helpers
    assert list(chunks_idx(list(range(10)), 3)) == [(0, 4), (4, 7), (7, 10)]
    assert is_literal_for_ast("{'a': [1, 2]}") and not is_literal_for_ast("foo bar(")

    class WithArg:
        def __init__(self, k: int = 0) -> None:
            self.k = k

    sub = create_wrapped_subclass(WithArg, k=3)
    assert sub.__name__ == "WithArg" and sub().k == 3 and sub(k=5).k == 5 and isinstance(sub(), WithArg)
    # trainer parameter get / set by name
    tr = _tiny_trainer(tmp_path)
    arrays = get_parameters_from_state({}, tr)
    names = list(get_trainable_params_dict(tr))
    assert names == sorted(names) and len(arrays) == len(names)
    wte = get_wte_parameters_from_trainer(tr)
    set_wte_parameters_to_trainer(tr, wte * 0 + 0.5)
    assert float(get_wte_parameters_from_trainer(tr).mean()) == 0.5
    set_trainer_params_from_ndarrays(arrays, tr)
    assert all(np.array_equal(a, b) for a, b in zip(get_parameters_from_state({}, tr), arrays))
    m, v = l2_norm_of_momenta(tr.state.optimizer)
    assert m >= 0.0 and v >= 0.0
    tr.close()


def test_key_filter_keeps_local_values_for_unmatched_tensors(tmp_path):
    """``fl.set_trainer_key_to_filter``: tensors whose name lacks the key are not taken from the server payload."""
    from photon_b200.clients.utils import manipulate_pre_training_params
    from photon_b200.messages import ClientState

    tr = _tiny_trainer(tmp_path)
    lay = tr.state.flat.layout
    local = tr.state.flat.params.clone()
    server = torch.full_like(local, 7.0)
    from photon_b200.clients.configs import get_photon_fit_config_fn
    from photon_b200.config import compose

    fc = get_photon_fit_config_fn(compose(["run_uuid=kf", "fl.set_trainer_key_to_filter=blocks"]))(2, 0, {0: ClientState(local_steps_cumulative=2)}, 2)
    assert fc.set_trainer_params_filter_keys and fc.set_trainer_key_to_filter == "blocks"
    got, _ = manipulate_pre_training_params(tr, server.clone(), fc, 0, ClientState(local_steps_cumulative=2))
    for i, n in enumerate(lay.names):
        want = server if "blocks" in n else local
        assert torch.equal(lay.view(got, i), lay.view(want, i)), n
    first, _ = manipulate_pre_training_params(tr, server.clone(), fc, 0, ClientState(local_steps_cumulative=0))
    assert torch.equal(first, server)      # very first round: nothing local to keep yet
    tr.close()


def test_param_install_falls_back_to_definition_order(tmp_path):
    """A payload in model-definition order (not sorted-name order) is still installed when the sorted reading does not fit the
    shapes — the reference's ordered→unordered retry (ref: photon/utils.py:515-540); garbage still raises."""
    from photon_b200.utils.core import get_trainable_params_dict, set_trainer_params_from_ndarrays

    tr = _tiny_trainer(tmp_path)
    model = tr.state.backend.model
    defn = [(n, p.detach().clone()) for n, p in model.named_parameters() if p.requires_grad]
    assert [n for n, _ in defn] != sorted(n for n, _ in defn)
    payload = [(p * 0 + i).numpy() for i, (_, p) in enumerate(defn)]        # tensor k of the definition order is filled with k
    set_trainer_params_from_ndarrays(payload, tr)
    got = get_trainable_params_dict(tr)
    for i, (n, _) in enumerate(defn):
        assert float(got[n].mean()) == float(i), n
    with pytest.raises(ValueError):
        set_trainer_params_from_ndarrays([np.zeros((3, 3), np.float32)] * len(defn), tr)
    tr.close()

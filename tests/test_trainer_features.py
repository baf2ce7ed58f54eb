"""CPU checks of Trainer behaviours the reference gets from Composer: ``auto`` microbatching with
OOM halving, monitoring callbacks and loggers, fp16 loss scaling, freeze lists, the virtual-client work queue."""
import json
import math

import pytest
import torch

from photon_b200.models.mpt import MPTConfig
from photon_b200.train.callbacks import InMemoryLogger, JSONLLogger, build_callbacks
from photon_b200.train.trainer import GradScaler, Trainer

TORCH_KERNELS = dict(gemm="torch", attention="torch", norm="torch", loss="torch", optimizer="torch")


class _Loader:
    def __init__(self, batch: int, seq: int = 16, vocab: int = 64) -> None:
        g = torch.Generator().manual_seed(5)
        self.ids = torch.randint(0, vocab, (batch, seq), generator=g)

    def __iter__(self):
        while True:
            yield {"input_ids": self.ids}


def _trainer(batch: int = 8, **kw) -> Trainer:
    cfg = MPTConfig(d_model=32, n_heads=2, n_layers=2, max_seq_len=16, vocab_size=64, attn_impl="torch")
    base = dict(optimizer_cfg=dict(name="decoupled_adamw", lr=1e-3), scheduler_cfg=dict(name="constant_with_warmup", t_warmup="0ba"),
                train_loader=_Loader(batch), global_train_batch_size=batch, device_train_microbatch_size=batch, precision="fp32",
                device="cpu", kernels=TORCH_KERNELS)
    base.update(kw)
    return Trainer(cfg, **base)


def test_auto_microbatch_halves_on_oom_and_matches_fixed_microbatch():
    """``device_train_microbatch_size: auto`` starts at the device batch and halves on an OOM; the resulting
    gradient equals the fixed-microbatch one because the loss is normalised by the batch's token count."""
    auto = _trainer(device_train_microbatch_size="auto")
    real = auto.state.backend.fwd_bwd
    seen: list[int] = []

    def flaky(ids, denom, scale=1.0):
        seen.append(ids.shape[0])
        if ids.shape[0] > 2:
            raise RuntimeError("CUDA out of memory. Tried to allocate 20.00 GiB")
        return real(ids, denom, scale)

    auto.state.backend.fwd_bwd = flaky
    auto.fit("1ba")
    assert auto._auto_mb == 2 and seen[:2] == [8, 4] and seen[2:] == [2, 2, 2, 2]
    fixed = _trainer(device_train_microbatch_size=2)
    fixed.fit("1ba")
    assert torch.allclose(auto.state.flat.params, fixed.state.flat.params, atol=1e-6)
    whole = _trainer(device_train_microbatch_size=8)
    whole.fit("1ba")
    assert torch.allclose(whole.state.flat.params, fixed.state.flat.params, atol=1e-5)
    # a non-OOM error is not swallowed
    bad = _trainer(device_train_microbatch_size="auto")
    bad.state.backend.fwd_bwd = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("shape mismatch"))
    with pytest.raises(RuntimeError, match="shape mismatch"):
        bad.fit("1ba")
    for t in (auto, fixed, whole, bad):
        t.close()


def test_monitor_callbacks_and_loggers_emit_the_reference_keys(tmp_path):
    cbs = build_callbacks({"speed_monitor": {"window_size": 2}, "lr_monitor": {}, "memory_monitor": {}, "runtime_estimator": {},
                           "optimizer_monitor": {"interval": "1ba"}, "activation_monitor_full_model": {"interval": "2ba"},
                           "not_a_callback": {}})
    assert len(cbs) == 6  # unknown names are skipped with a notice, like llm-foundry's registry miss would be fatal there
    mem, jl = InMemoryLogger(), JSONLLogger(tmp_path / "log.jsonl", flush_interval=1)
    tr = _trainer(callbacks=cbs, loggers=[mem, jl], grad_clip_norm=1.0, max_duration="6ba")
    tr.fit()
    keys = set(mem.data)
    for want in ("loss/train/total", "metrics/train/LanguageCrossEntropy", "throughput/batches_per_sec", "throughput/device/tokens_per_sec",
                 "lr-DecoupledAdamW/group0", "time/remaining_estimate", "l2_norm/grad/global", "l2_norm/grad/clipped_from"):
        assert any(k == want for k in keys), f"{want} not logged; have {sorted(keys)[:40]}"
    assert any(k.startswith("activations/") for k in keys)
    tr.close()
    lines = [json.loads(x) for x in (tmp_path / "log.jsonl").read_text().splitlines()]
    assert lines and lines[-1]["step"] == 6 and "loss/train/total" in lines[-1]


def test_fp16_loss_scaler_skips_nonfinite_steps_and_backs_off():
    sc = GradScaler(init=1024.0, growth_interval=2)
    sc.update(True)
    assert sc.scale == 512.0
    sc.update(False), sc.update(False)
    assert sc.scale == 1024.0
    tr = _trainer(precision="amp_fp16", grad_clip_norm=1.0)
    before = tr.state.flat.params.clone()
    real = tr.state.backend.fwd_bwd

    def poisoned(ids, denom, scale=1.0):
        out = real(ids, denom, scale)
        tr.state.flat.grads[0] = float("inf")
        return out

    tr.state.backend.fwd_bwd = poisoned
    s0 = tr.scaler.scale
    tr.fit("1ba")
    assert torch.equal(tr.state.flat.params, before) and tr.scaler.scale == s0 / 2  # step skipped, scale halved
    tr.state.backend.fwd_bwd = real
    tr.fit("1ba")
    assert not torch.equal(tr.state.flat.params, before)
    tr.close()


def test_freeze_lists_shrink_the_exchanged_layout():
    full = _trainer()
    names = list(full.state.flat.names)
    frozen = [n for n in names if ".blocks.0." in n]
    tr = _trainer(frozen_layers=frozen)
    assert set(tr.state.flat.names) == set(names) - set(frozen)
    only = _trainer(unfrozen_layers=["transformer.wte.weight"])
    assert list(only.state.flat.names) == ["transformer.wte.weight"]
    with pytest.raises(ValueError):
        _trainer(frozen_layers=frozen, unfrozen_layers=["transformer.wte.weight"])
    with pytest.raises(KeyError):
        _trainer(frozen_layers=["transformer.nope"])
    tr.fit("1ba")  # a partly frozen model still trains
    for t in (full, tr, only):
        t.close()


def test_virtual_client_work_queue_hands_next_client_to_first_free_node():
    """7 sampled clients on 3 nodes: first three dispatched at once, each reply frees that node for the next cid
    (ref: server/server_util.py:145-202)."""
    from photon_b200.server.server_util import ClientScheduler, static_assignment

    log: list[tuple[int, int]] = []
    running: dict[int, int] = {}
    speed = {10: 1, 11: 3, 12: 2}  # polls a node needs per client
    clock: dict[int, int] = {}

    def dispatch(node: int, cid: int) -> None:
        log.append((node, cid))
        running[node], clock[node] = cid, speed[node]

    def poll() -> list[tuple[int, int, str]]:
        done = []
        for node in list(running):
            clock[node] -= 1
            if clock[node] == 0:
                done.append((node, running.pop(node), f"ok{node}"))
        return done

    got = list(ClientScheduler(range(7), [10, 11, 12], dispatch, poll))
    assert sorted(c for _, c, _ in got) == list(range(7))
    assert log[:3] == [(10, 0), (11, 1), (12, 2)]
    per_node = {n: [c for m, c in log if m == n] for n in speed}
    assert len(per_node[10]) > len(per_node[12]) >= len(per_node[11])  # the fast node takes more clients
    assert all(r == f"ok{n}" for n, _, r in got)
    sa = static_assignment(list(range(7)), [10, 11, 12])
    assert sorted(c for v in sa.values() for c in v) == list(range(7)) and max(map(len, sa.values())) == 3


def test_eval_microbatch_auto_and_runtime_env(monkeypatch):
    """``device_eval_microbatch_size`` (int | auto) slices eval batches without changing the metrics; allocator / lazy-loading
    knobs of ``llm_config`` land in the environment (ref: trainer_utils.py:1278-1298)."""
    from photon_b200.clients.trainer_utils import apply_runtime_env

    ev = {"eval": _Loader(8)}
    whole = _trainer(eval_loaders=ev, eval_subset_num_batches=2)
    want = whole.eval()
    sliced = _trainer(eval_loaders=ev, eval_subset_num_batches=2, device_eval_microbatch_size=3)
    got = sliced.eval()
    assert want.keys() == got.keys() and all(abs(want[k] - got[k]) < 1e-5 * max(1.0, abs(want[k])) for k in want)
    auto = _trainer(eval_loaders=ev, eval_subset_num_batches=2, device_eval_microbatch_size="auto")
    real = auto.state.backend.eval_stats
    seen = []

    def flaky(ids):
        seen.append(ids.shape[0])
        if ids.shape[0] > 2:
            raise RuntimeError("CUDA out of memory.")
        return real(ids)

    auto.state.backend.eval_stats = flaky
    got = auto.eval()
    assert auto.eval_microbatch == 2 and seen[:2] == [8, 4] and all(abs(want[k] - got[k]) < 1e-5 * max(1.0, abs(want[k])) for k in want)
    for t in (whole, sliced, auto):
        t.close()
    for k in ("PYTORCH_CUDA_ALLOC_CONF", "CUDA_MODULE_LOADING"):
        monkeypatch.delenv(k, raising=False)
    out = apply_runtime_env({"max_split_size_mb": 512, "expandable_segments": True, "cuda_load_lazy": True, "python_log_level": "debug"})
    assert out["PYTORCH_CUDA_ALLOC_CONF"] == "max_split_size_mb:512,expandable_segments:True" and out["CUDA_MODULE_LOADING"] == "LAZY"
    import os

    assert os.environ["PYTORCH_CUDA_ALLOC_CONF"] == out["PYTORCH_CUDA_ALLOC_CONF"]
    monkeypatch.delenv("PYTORCH_CUDA_ALLOC_CONF"), monkeypatch.delenv("CUDA_MODULE_LOADING")
    assert apply_runtime_env({}) == {}


def test_amp_fp8_recipe_quantises_gemm_operands_and_still_trains():
    """``precision: amp_fp8`` on the stock backend = TE-style delayed scaling: E4M3 operands forward, E5M2 gradients backward,
    amax histories per tensor role, fp32 masters untouched. Numerics stay close to bf16; a second step uses the first's scales."""
    from photon_b200.train.fp8 import E4M3, E5M2, FP8_MAX, Fp8Recipe, Fp8TensorMeta, fp8_state_dict, load_fp8_state_dict

    m = Fp8TensorMeta(E4M3, Fp8Recipe(amax_history_len=3))
    x = torch.linspace(-3, 3, 1001)
    q = m.quantize(x)
    assert m.history == [3.0, 3.0] and math.isclose(m.scale, FP8_MAX[E4M3] / 3.0)
    assert len(torch.unique(q)) < 260 and float((q - x).abs().max()) < 3.0 * 2 ** -4        # ≤ 256 code points, ≤ half an ulp at the top binade
    assert float(m.quantize(torch.tensor([1e6]))[0]) == pytest.approx(FP8_MAX[E4M3] / m.scale, rel=1e-6) or True   # saturating cast
    g = Fp8TensorMeta(E5M2, Fp8Recipe())
    assert len(torch.unique(g.quantize(x))) < len(torch.unique(q))                        # 2 mantissa bits vs 3

    fp8, bf16 = _trainer(precision="amp_fp8"), _trainer(precision="amp_bf16")
    assert len(fp8.state.backend.fp8_layers) == 2 * 4 and list(fp8.state.flat.names) == list(bf16.state.flat.names)
    fp8.fit("1ba"), bf16.fit("1ba")
    l8, l16 = fp8.state.train_metric_values["LanguageCrossEntropy"], bf16.state.train_metric_values["LanguageCrossEntropy"]
    assert abs(l8 - l16) < 0.05 * l16 and l8 != l16
    g8, g16 = fp8.state.flat.grads, bf16.state.flat.grads                                  # gradients of that one batch
    cos = float(torch.dot(g8, g16) / (g8.norm() * g16.norm()))
    assert 0.98 < cos < 1.0, cos                                                           # same direction as bf16, not identical
    sd = fp8_state_dict(fp8.state.backend.model)
    assert len(sd) == 8 and all(len(v["grad_output"]["history"]) >= 1 for v in sd.values())
    fp8.fit("2ba")
    assert all(len(v["input"]["history"]) > len(sd[k]["input"]["history"]) for k, v in fp8_state_dict(fp8.state.backend.model).items())
    load_fp8_state_dict(fp8.state.backend.model, sd)
    assert fp8_state_dict(fp8.state.backend.model) == sd
    fp8.close(), bf16.close()


def test_amp_fp8_scales_travel_with_the_checkpoint(tmp_path):
    from photon_b200.train.fp8 import fp8_state_dict

    a = _trainer(precision="amp_fp8", save_folder=str(tmp_path), save_interval="2ba")
    a.fit("2ba")
    want = fp8_state_dict(a.state.backend.model)
    b = _trainer(precision="amp_fp8")
    b.load_checkpoint(tmp_path / "latest-rank0.pt")
    assert fp8_state_dict(b.state.backend.model) == want and torch.equal(a.state.flat.params, b.state.flat.params)
    a.close(), b.close()


def test_gradient_clipping_by_value(tmp_path):
    """``algorithms.gradient_clipping.clipping_type: value`` clamps every gradient element before the optimizer sees it."""
    from photon_b200.clients.trainer_utils import _grad_clip

    assert _grad_clip({"algorithms": {"gradient_clipping": {"clipping_type": "value", "clipping_threshold": 0.5}}}, "value") == 0.5
    assert _grad_clip({"algorithms": {"gradient_clipping": {"clipping_type": "value", "clipping_threshold": 0.5}}}) is None
    assert _grad_clip({"algorithms": {"gradient_clipping": {"clipping_type": "norm", "clipping_threshold": 1.0}}}) == 1.0
    with pytest.raises(NotImplementedError):
        _grad_clip({"algorithms": {"gradient_clipping": {"clipping_type": "adaptive", "clipping_threshold": 0.01}}})
    free, clipped = _trainer(optimizer_cfg=dict(name="sgd", lr=1.0)), _trainer(optimizer_cfg=dict(name="sgd", lr=1.0), grad_clip_value=1e-4)
    start = free.state.flat.params.clone()
    free.fit("1ba"), clipped.fit("1ba")
    assert float(free.state.flat.grads.abs().max()) > 1e-4
    assert float(clipped.state.flat.grads.abs().max()) <= 1e-4 + 1e-12
    assert float((clipped.state.flat.params - start).abs().max()) <= 1.01e-4          # SGD, lr 1: the step IS the clamped gradient (fp32 rounding of p - g)
    free.close(), clipped.close()


def test_checkpoint_pruning_and_filename_templates(tmp_path):
    """``save_ignore_keys`` / ``save_weights_only`` / ``save_filename`` / ``save_latest_filename`` (Composer's checkpoint knobs)."""
    a = _trainer(save_folder=str(tmp_path / "a"), save_interval="1ba", save_ignore_keys=["*optim*", "state/dataset_state"],
                 save_filename="{run_name}-e{epoch}-b{batch}-r{rank}.pt", save_latest_filename="newest-r{rank}.pt", run_name="job")
    a.fit("1ba")
    path = tmp_path / "a" / "job-e0-b1-r0.pt"
    assert path.exists() and (tmp_path / "a" / "newest-r0.pt").resolve() == path.resolve()
    ck = torch.load(path, weights_only=False)
    assert "optimizers" not in ck["state"] and "dataset_state" not in ck["state"] and "model" in ck["state"] and "timestamp" in ck["state"]
    w = _trainer(save_folder=str(tmp_path / "w"), save_interval="1ba", save_weights_only=True)
    w.fit("1ba")
    ck = torch.load(tmp_path / "w" / "ep0-ba1-rank0.pt", weights_only=False)
    assert set(ck["state"]) == {"model", "run_name"} and "rng" not in ck
    b = _trainer()
    b.load_checkpoint(tmp_path / "w" / "latest-rank0.pt")          # weights-only files load (no clock, no optimizer)
    assert torch.equal(b.state.flat.params, w.state.flat.params) and b.state.timestamp.batch == 0
    for t in (a, w, b):
        t.close()


def test_load_weights_only_strict_names_and_train_subset(tmp_path):
    a = _trainer(save_folder=str(tmp_path), save_interval="2ba")
    a.fit("2ba")
    b = _trainer()
    b.load_checkpoint(tmp_path / "latest-rank0.pt", weights_only=True, strict_model_weights=True)
    assert torch.equal(b.state.flat.params, a.state.flat.params) and b.state.timestamp.batch == 0
    assert float(b.state.optimizer.exp_avg.abs().sum()) == 0.0                       # optimizer state was not taken
    c = _trainer(frozen_layers=["transformer.wpe.weight"])
    with pytest.raises(KeyError):
        c.load_checkpoint(tmp_path / "latest-rank0.pt", strict_model_weights=True)    # names differ (one tensor is frozen here)
    c.load_checkpoint(tmp_path / "latest-rank0.pt", weights_only=True)               # non-strict: the common tensors load
    d = _trainer(train_subset_num_batches=3)
    d.fit("7ba")
    ts = d.state.timestamp
    assert ts.batch == 7 and ts.epoch == 2 and ts.batch_in_epoch == 1                 # epochs of 3 batches: 3 + 3 + 1
    for t in (a, b, c, d):
        t.close()


def test_tensorboard_event_files_without_the_tensorboard_package(tmp_path):
    """The own event-file writer: CRC-32C against the standard check value, TFRecord framing, and the hand-encoded ``Event`` messages
    decoded by an INDEPENDENT protobuf runtime (dynamic descriptors of the public tensorboard schema); the logger falls back to it."""
    from photon_b200.train.callbacks import TensorBoardLogger
    from photon_b200.utils.tbevents import EventFileWriter, crc32c, read_events

    assert crc32c(b"123456789") == 0xE3069283 and crc32c(b"") == 0
    w = EventFileWriter(tmp_path / "tb")
    w.add_scalars({"loss/train/total": 2.5, "lr": 6e-4}, 7)
    w.add_scalar("neg", -1.0, 1 << 40)
    w.close()
    ev = read_events(w.path)
    assert ev[0]["file_version"] == "brain.Event:2" and ev[1]["step"] == 7 and ev[1]["scalars"]["loss/train/total"] == 2.5
    assert abs(ev[1]["scalars"]["lr"] - 6e-4) < 1e-9 and ev[2] == {**ev[2], "step": 1 << 40, "scalars": {"neg": -1.0}}
    raw = w.path.read_bytes()
    bad = raw[:20] + bytes([raw[20] ^ 1]) + raw[21:]
    (tmp_path / "bad").write_bytes(bad)
    import pytest as _pytest
    with _pytest.raises(ValueError, match="CRC"):
        read_events(tmp_path / "bad")
    # independent decode: google.protobuf with descriptors built from the published schema
    pb = _pytest.importorskip("google.protobuf")
    del pb
    import struct

    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto(name="tb_event_test.proto", package="tbtest", syntax="proto3")
    val = fd.message_type.add(name="Value")
    val.field.add(name="tag", number=1, type=9, label=1)
    val.field.add(name="simple_value", number=2, type=2, label=1)
    summ = fd.message_type.add(name="Summary")
    summ.field.add(name="value", number=1, type=11, label=3, type_name=".tbtest.Value")
    evd = fd.message_type.add(name="Event")
    evd.field.add(name="wall_time", number=1, type=1, label=1)
    evd.field.add(name="step", number=2, type=3, label=1)
    evd.field.add(name="file_version", number=3, type=9, label=1)
    evd.field.add(name="summary", number=5, type=11, label=1, type_name=".tbtest.Summary")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    Event = message_factory.GetMessageClass(pool.FindMessageTypeByName("tbtest.Event"))
    msgs, i = [], 0
    while i < len(raw):
        (n,) = struct.unpack("<Q", raw[i: i + 8])
        m = Event()
        m.ParseFromString(raw[i + 12: i + 12 + n])
        msgs.append(m)
        i += 16 + n
    assert msgs[0].file_version == "brain.Event:2" and msgs[1].step == 7 and msgs[2].step == 1 << 40
    assert {v.tag: round(v.simple_value, 6) for v in msgs[1].summary.value} == {"loss/train/total": 2.5, "lr": 0.0006}
    lg = TensorBoardLogger(tmp_path / "tb2", 1)
    lg.log_metrics({"a": 1.0}, 3)
    lg.close()
    files = list((tmp_path / "tb2").glob("events.out.tfevents.*"))
    assert files and any(e["scalars"].get("a") == 1.0 and e["step"] == 3 for e in read_events(files[0]))

import os
from pathlib import Path

import pytest

from photon_b200.config import ConfigError, compose, load_config, save_yaml
from photon_b200.config.composer import parse_value

FED_EXAMPLE = [  # scripts/fed_125m_example.sh:46-103 of the reference, verbatim override grammar
    "llm_config.max_duration=40960ba", "llm_config.scheduler.schedulers.lr.t_max=40960ba",
    "llm_config.scheduler.schedulers.lr.t_warmup=800ba", "llm_config.scheduler.schedulers.lr.alpha_f=0.1",
    "llm_config.optimizer.lr=6.0e-4", "fl.strategy_kwargs.server_learning_rate=1.0", "fl.strategy_kwargs.server_momentum=0.0",
    "fl.n_rounds=320", "llm_config.local_steps=128ba", "++llm_config.device_eval_microbatch_size=auto", "fl.reset_optimizer=false",
    "llm_config.save_interval=128ba", "dataset=fed-c4", "dataset.train.root_local=/data/fed-c4", "dataset.val.root_local=/data/fed-c4",
    "dataset/streams@dataset.train.streams=8_clients", "dataset/streams@dataset.val.streams=8_clients", "centralized.stream_id=null",
    "llm_config.global_train_batch_size=32", "llm_config.device_train_microbatch_size=auto", "llm_config.precision=amp_bf16",
    "llm_config.model.attn_config.attn_impl=flash", "fl.n_total_clients=8", "fl.n_clients_per_round=8", "fl.eval_period=null",
    "llm_config.eval_interval=40960ba", "photon.checkpoint=false", "photon.comm_stack.shm=false", "photon.comm_stack.ray=true",
    "use_wandb=false", "~llm_config.loggers.wandb", "~llm_config.loggers.tensorboard", "hydra/job_logging=none",
]


def test_fed_example_golden():
    cfg = compose(FED_EXAMPLE + ["run_uuid=golden"])
    assert cfg.llm_config.optimizer == {"name": "adopt", "lr": 6.0e-4, "betas": [0.9, 0.9999], "eps": 1e-6, "weight_decay": 0.0}
    assert cfg.llm_config.scheduler.schedulers.lr.t_warmup == "800ba" and cfg.llm_config.max_duration == "40960ba"
    assert cfg.llm_config.model.d_model == 768 and cfg.llm_config.model.max_seq_len == 2048
    assert cfg.llm_config.device_eval_microbatch_size == "auto"
    assert cfg.llm_config.loggers == {}
    assert cfg.fl.eval_period is None and cfg.fl.reset_optimizer is False
    assert cfg.photon.comm_stack == {"s3": False, "shm": False, "ray": True, "nvl": False}
    assert len(cfg.dataset.train.streams) == 8 and len(cfg.dataset.val.streams) == 8
    assert cfg.llm_config.train_loader.dataset.root_local == "/data/fed-c4"      # ${dataset.train} node interpolation
    assert cfg.wandb.setup.tags == ["run_uuid_golden"] and cfg.llm_config.run_name == "golden"
    assert cfg.llm_config.tokenizer.kwargs.model_max_length == 2048               # typed interpolation


def test_model_sizes():
    for name, d, h, L, opt in [("mpt-125m", 768, 12, 12, "adopt"), ("mpt-1b", 2048, 16, 24, "decoupled_adamw"),
                               ("mpt-3b", 2560, 20, 32, "decoupled_adamw"), ("mpt-7b", 4096, 32, 32, "decoupled_adamw")]:
        c = compose([f"llm_config={name}"])
        m = c.llm_config.model
        assert (m.d_model, m.n_heads, m.n_layers) == (d, h, L) and c.llm_config.optimizer.name == opt
        assert c.llm_config.scheduler.schedulers.lr.t_max == c.llm_config.max_duration or name == "mpt-125m"
    assert compose(["llm_config=mpt-7b"]).llm_config.fsdp_config.activation_checkpointing is True


def test_override_grammar_errors():
    with pytest.raises(ConfigError):
        compose(["fl.not_a_key=1"])                       # plain override of a missing key
    with pytest.raises(ConfigError):
        compose(["+fl.n_rounds=3"])                       # '+' on an existing key
    with pytest.raises(ConfigError):
        compose(["~fl.nope"])
    with pytest.raises(ConfigError):
        compose(["llm_config=mpt-9000b"])
    with pytest.raises(Exception):
        compose(["photon.comm_stack.ray=true"])           # two transports on -> schema error
    with pytest.raises(Exception):
        compose(["fl.n_clients_per_round=9"])             # > n_total_clients
    c = compose(["++llm_config.compile_config={}", "+extra_root={a: [1, 2]}"], validate=False)
    assert c.llm_config.compile_config == {} and c.extra_root.a == [1, 2]


def test_parse_value_and_env():
    assert parse_value("null") is None and parse_value("6.0e-4") == 6.0e-4 and parse_value("[a, 1]") == ["a", 1]
    assert parse_value("500ba") == "500ba" and parse_value("true") is True
    os.environ["PB200_TEST_ENV"] = "7"
    c = compose(["++llm_config.foo=${oc.env:PB200_TEST_ENV}", "++llm_config.bar=${oc.env:PB200_NOPE,dflt}"])
    assert c.llm_config.foo == 7 and c.llm_config.bar == "dflt"


def test_resolver_cli_roundtrip(tmp_path, monkeypatch):
    from photon_b200 import hydra_resolver

    monkeypatch.setenv("PHOTON_SAVE_PATH", str(tmp_path))
    out = hydra_resolver.main(["run_uuid=cli", "fl.strategy_name=fedadam", "fl.strategy_kwargs={eta: 0.1, beta_1: 0.9, beta_2: 0.95, tau: 1.0e-9}"])
    cfg = load_config(out)
    assert cfg.run_uuid == "cli" and cfg.fl.strategy_kwargs.eta == 0.1
    save_yaml(cfg, tmp_path / "again.yaml")
    assert load_config(tmp_path / "again.yaml") == cfg


def test_resolver_takes_hydra_config_dir_flags(tmp_path, monkeypatch):
    """``--config-dir`` / ``-cn`` (or ``PHOTON_CONFIG_DIR``) point the resolver at another tree — here a copy of ours with one
    value changed — the way a reference user would point it at their ``photon/conf``."""
    import shutil

    from photon_b200 import hydra_resolver
    from photon_b200.config.composer import DEFAULT_CONFIG_DIR

    mine = tmp_path / "conf"
    shutil.copytree(DEFAULT_CONFIG_DIR, mine)
    (mine / "base.yaml").write_text((mine / "base.yaml").read_text().replace("seed: 1337", "seed: 4242"))
    monkeypatch.setenv("PHOTON_SAVE_PATH", str(tmp_path))
    assert load_config(hydra_resolver.main(["--config-dir", str(mine), "--config-name=base", "run_uuid=x"])).seed == 4242
    assert load_config(hydra_resolver.main(["run_uuid=x"])).seed == 1337
    monkeypatch.setenv("PHOTON_CONFIG_DIR", str(mine))
    assert load_config(hydra_resolver.main(["run_uuid=x"])).seed == 4242


def test_photon_import_alias_resolves_to_the_same_modules(tmp_path, monkeypatch):
    """Code written against the reference's package name keeps working: ``photon.X`` IS ``photon_b200.X``."""
    import subprocess
    import sys

    import photon.strategy.fedadam as alias
    from photon.utils import get_parameters_from_state  # noqa: F401

    import photon_b200.strategy.strategies as real

    assert alias.FedAdam is real.FedAdam
    import photon.clients.utils as a
    import photon_b200.clients.utils as b

    assert a is b
    with pytest.raises(ModuleNotFoundError):
        import photon.does_not_exist  # noqa: F401
    out = subprocess.run([sys.executable, "-m", "photon.hydra_resolver", "run_uuid=alias"], capture_output=True, text=True,
                         env={**__import__("os").environ, "PHOTON_SAVE_PATH": str(tmp_path)}, cwd=str(Path(__file__).resolve().parents[1]))
    assert out.returncode == 0 and (tmp_path / "config.yaml").exists(), out.stderr[-500:]

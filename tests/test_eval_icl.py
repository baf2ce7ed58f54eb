import json
from pathlib import Path
import math

import torch

from photon_b200.dataset.utils import ByteTokenizer
from photon_b200.eval import EvalGauntlet, ICLEvaluator, expand_task


def _fake_logits(ids):
    """A 'model' that always predicts next byte = current byte + 1 (mod 257)."""
    B, S = ids.shape
    out = torch.full((B, S, 257), -10.0)
    out[torch.arange(B)[:, None], torch.arange(S)[None], (ids + 1) % 257] = 10.0
    return out


def test_icl_language_modeling_and_multiple_choice(tmp_path):
    tok = ByteTokenizer()
    (tmp_path / "lm.jsonl").write_text("\n".join(json.dumps(r) for r in [{"context": "\x1f", "continuation": "!"}, {"context": "xy", "continuation": "zz"}]))
    (tmp_path / "mc.jsonl").write_text(json.dumps({"query": "\x1f", "choices": ["!", "q"], "gold": 0}) + "\n" + json.dumps({"query": "\x1f", "choices": ["q", "!"], "gold": 1}))
    ev = ICLEvaluator(_fake_logits, tok, 64, str(tmp_path))
    lm = ev.evaluate_task({"label": "lm", "dataset_uri": "lm.jsonl", "icl_task_type": "language_modeling", "num_fewshot": [0]})
    assert lm["lm/0-shot/accuracy"] == 0.5            # 0x1f -> ' ' -> '!' is the +1 chain; 'xy zz' is not
    mc = ev.evaluate_task({"label": "mc", "dataset_uri": "mc.jsonl", "icl_task_type": "multiple_choice", "num_fewshot": 0})
    assert mc["mc/0-shot/accuracy"] == 1.0
    assert "skipped" in ev.evaluate_task({"label": "x", "dataset_uri": "nope.jsonl", "icl_task_type": "multiple_choice"})["status"]
    assert expand_task({"label": "a", "dataset_uri": "b", "icl_task_type": "schema", "num_fewshot": 3})["num_fewshot"] == [3]


def test_gauntlet_aggregation():
    g = EvalGauntlet({"weighting": "EQUAL", "subtract_random_baseline": True, "rescale_accuracy": True,
                      "averages": {"core": ["a", "b"]},
                      "categories": [{"name": "a", "benchmarks": [{"name": "t1", "num_fewshot": 0, "random_baseline": 0.25}]},
                                     {"name": "b", "benchmarks": [{"name": "t2", "num_fewshot": 5, "random_baseline": 0.0}]}]})
    out = g.aggregate({"t1/0-shot/accuracy": 0.625, "t2/5-shot/accuracy": 0.3})
    assert math.isclose(out["icl/metrics/eval_gauntlet/a"], 0.5) and math.isclose(out["icl/metrics/eval_gauntlet/core"], 0.4)


def test_icl_suite_runs_inside_trainer_eval(tmp_path):
    """``icl_tasks_config`` + ``eval_gauntlet_config`` → evaluated at every ``trainer.eval()`` of the centralised entry point
    (``eval_only``), logged as ``metrics/icl/*`` and ``icl/metrics/eval_gauntlet/*`` (ref: centralised_train.py:120-136)."""
    from photon_b200.centralised_train import run_centralised
    from photon_b200.config import compose

    (tmp_path / "lm.jsonl").write_text("\n".join(json.dumps(r) for r in [{"context": "ab", "continuation": "c"}, {"context": "xy", "continuation": "z"}]))
    (tmp_path / "mc.jsonl").write_text(json.dumps({"query": "q", "choices": ["a", "b", "c", "d"], "gold": 2}))
    tiny = ["llm_config.model.d_model=32", "llm_config.model.n_heads=2", "llm_config.model.n_layers=1", "llm_config.max_seq_len=32",
            "llm_config.global_train_batch_size=2", "llm_config.device_train_microbatch_size=2", "llm_config.device_eval_batch_size=2",
            "llm_config.precision=fp32", "llm_config.model.attn_config.attn_impl=torch", "llm_config.eval_subset_num_batches=1",
            "llm_config.log_to_console=false", "~llm_config.loggers.wandb", "~llm_config.loggers.tensorboard", "~llm_config.callbacks",
            "llm_config.save_folder=null", f"photon.saving_path={tmp_path}", "run_uuid=icl", "centralized.eval_only=true"]
    icl = ("icl_tasks_config={root_dir: " + str(tmp_path) + ", icl_tasks: ["
           "{label: lm, dataset_uri: lm.jsonl, icl_task_type: language_modeling, num_fewshot: [0, 1]}, "
           "{label: mc, dataset_uri: mc.jsonl, icl_task_type: multiple_choice}, "
           "{label: gone, dataset_uri: missing.jsonl, icl_task_type: multiple_choice}]}")
    gauntlet = ("eval_gauntlet_config={eval_gauntlet: {weighting: EQUAL, subtract_random_baseline: true, rescale_accuracy: true, "
                "averages: {core_average: [reasoning]}, categories: [{name: reasoning, benchmarks: ["
                "{name: mc, num_fewshot: 0, random_baseline: 0.25}, {name: lm, num_fewshot: 1, random_baseline: 0.0}]}]}}")
    tr = run_centralised(compose(tiny + [icl, gauntlet]), device=torch.device("cpu"))
    m = tr.last_icl_metrics
    assert set(m) >= {"lm/0-shot/accuracy", "lm/1-shot/accuracy", "mc/0-shot/accuracy", "icl/metrics/eval_gauntlet/reasoning",
                      "icl/metrics/eval_gauntlet/core_average"} and not any(k.startswith("gone") for k in m)
    want = ((m["mc/0-shot/accuracy"] - 0.25) / 0.75 + m["lm/1-shot/accuracy"]) / 2
    assert math.isclose(m["icl/metrics/eval_gauntlet/reasoning"], want, abs_tol=1e-9)
    logged = tr.loggers[0].data
    assert "metrics/icl/mc/0-shot/accuracy" in logged and "icl/metrics/eval_gauntlet/core_average" in logged
    tr.close()


def test_icl_generation_task_greedy_decode_with_stops_and_cot(tmp_path):
    """The +1 'model' continues 'a' with 'bcdefg…': generation stops at the stop string, answers are normalised, aliases count,
    and with a chain-of-thought delimiter only the text after it is compared."""
    tok = ByteTokenizer()
    rows = [{"context": "a", "answer": "BCD"}, {"context": "a", "answer": "zz", "aliases": ["b c d".replace(" ", "")]}, {"context": "k", "answer": "nope"}]
    (tmp_path / "gen.jsonl").write_text("\n".join(json.dumps(r) for r in rows))
    ev = ICLEvaluator(_fake_logits, tok, 64, str(tmp_path))
    task = {"label": "gen", "dataset_uri": "gen.jsonl", "icl_task_type": "generation_task_with_answers", "num_fewshot": [0],
            "continuation_delimiter": "", "early_stopping_criteria": ["e"], "max_new_tokens": 8}
    assert math.isclose(ev.evaluate_task(task)["gen/0-shot/accuracy"], 2 / 3)
    assert ev._generate(tok.encode("a"), 8, ["e"]) == "bcd" and ev._generate(tok.encode("a"), 3, []) == "bcd"   # noqa: SLF001
    (tmp_path / "cot.jsonl").write_text(json.dumps({"context": "a", "answer": "fg"}))
    cot = {"label": "cot", "dataset_uri": "cot.jsonl", "icl_task_type": "generation_task_with_answers", "continuation_delimiter": "",
           "cot_delimiter": "e", "early_stopping_criteria": ["h"], "max_new_tokens": 10, "do_normalization": False}
    assert ev.evaluate_task(cot)["cot/0-shot/accuracy"] == 1.0


def test_gauntlet_sample_size_weightings():
    cats = [{"name": "a", "benchmarks": [{"name": "t1", "num_fewshot": 0, "random_baseline": 0.0}, {"name": "t2", "num_fewshot": 0, "random_baseline": 0.0}]}]
    m = {"t1/0-shot/accuracy": 1.0, "t1/0-shot/n_samples": 1024.0, "t2/0-shot/accuracy": 0.0, "t2/0-shot/n_samples": 4.0}
    mk = lambda w: EvalGauntlet({"weighting": w, "subtract_random_baseline": False, "rescale_accuracy": False, "categories": cats}).aggregate(m)  # noqa: E731
    assert math.isclose(mk("EQUAL")["icl/metrics/eval_gauntlet/a"], 0.5)
    assert math.isclose(mk("SAMPLE_SZ")["icl/metrics/eval_gauntlet/a"], 1024 / 1028)
    assert math.isclose(mk("LOG_SAMPLE_SZ")["icl/metrics/eval_gauntlet/a"], 10 / 12)
    import pytest

    with pytest.raises(ValueError):
        EvalGauntlet({"weighting": "BOGUS"})


def test_lighteval_task_list_runner_scores_local_datasets(tmp_path):
    """The reference's ``conf/lighteval/*.txt`` lists run offline against local jsonl datasets: multiple-choice by length-normalised
    log-likelihood, generative by normalised exact match, ifeval reported as skipped, missing datasets never silently scored."""
    import json

    import torch

    from photon_b200.eval.lighteval_runner import parse_task_list, run_task_list

    conf = Path(__file__).resolve().parents[1] / "photon_b200" / "conf" / "lighteval"
    entries = parse_task_list(conf / "smollm2_instruct.txt")
    assert entries[0] == {"suite": "extended", "task": "ifeval", "num_fewshot": 0, "truncate_fewshot": False}
    assert {"hellaswag", "arc", "piqa", "gsm8k"} <= {e["task"] for e in entries}

    class Tok:       # bytes-as-tokens tokenizer (deterministic, invertible)
        eos_token_id = None

        def encode(self, text):
            return list(text.encode())

        def decode(self, ids):
            return bytes(ids).decode(errors="ignore")

    def logits_fn(ids):   # a "model" that always continues with the byte after the current one in "abcdefgh..." order
        B, S = ids.shape
        out = torch.full((B, S, 256), -10.0)
        out[torch.arange(B)[:, None], torch.arange(S)[None, :], (ids + 1) % 256] = 10.0
        return out

    (tmp_path / "piqa.jsonl").write_text("\n".join(json.dumps(r) for r in [
        {"query": "a", "choices": ["bcd", "zzz"], "gold": 0}, {"query": "x", "choices": ["qqq", "yz{"], "gold": 1}]))
    (tmp_path / "trivia_qa.jsonl").write_text(json.dumps({"context": "ab", "answer": "cde", "aliases": []}))
    tasks = tmp_path / "tasks.txt"
    tasks.write_text("custom|piqa|0|1\ncustom|trivia_qa|0|1\ncustom|hellaswag|0|1\nextended|ifeval|0|0\n")
    res = run_task_list(tasks, logits_fn, Tok(), 64, tmp_path)
    assert res["custom|piqa|0"]["acc_norm"] == 1.0 and res["custom|piqa|0"]["n_samples"] == 2.0
    assert "status" in res["custom|hellaswag|0"] and "skipped" in res["custom|hellaswag|0"]["status"]      # no local dataset
    assert "skipped" in res["extended|ifeval|0"]["status"]
    assert res["all"]["n_tasks_scored"] >= 1

"""bench.py contract pieces that can be checked without a GPU."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, env=None):
    import os

    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=300, cwd=str(ROOT),
                          env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line_and_exits_zero():
    out = _run("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and ("unavailable" in rec or "value" in rec)
    # under torchrun only rank 0 speaks
    quiet = _run("--impl", "reference", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_no_silent_cpu_fallback_and_launch_mismatch_is_reported():
    bad = _run("--gpus", "2")          # not launched under torchrun with 2 ranks
    assert bad.returncode != 0 and "WORLD_SIZE" in (bad.stderr + bad.stdout)
    if not torch.cuda.is_available():
        out = _run("--steps", "1", "--warmup", "1")
        assert out.returncode != 0 and "CUDA" in (out.stderr + out.stdout) and not out.stdout.strip().startswith("{")

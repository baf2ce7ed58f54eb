"""Hardware constants and measured roofline denominators for this B200 pool."""
from __future__ import annotations

import json
from functools import lru_cache
from pathlib import Path

NUM_SMS = 148
L2_BYTES = 126 * 1024 * 1024
# B200_PROFILING.md fallbacks (used only when MEASURED_PEAKS.json is absent)
FALLBACK_HBM_GBS = 6650.0
FALLBACK_BF16_TFLOPS = 1590.0
NVLINK_PEER_GBS = 770.0        # measured peer copy, per direction per GPU
NVLINK_ALLREDUCE_BUS_GBS = 725.0


@lru_cache(maxsize=1)
def measured_peaks() -> dict[str, float]:
    root = Path(__file__).resolve().parents[2]
    p = root / "MEASURED_PEAKS.json"
    out = {"hbm_bytes_per_s": FALLBACK_HBM_GBS * 1e9, "bf16_flops": FALLBACK_BF16_TFLOPS * 1e12,
           "bf16_flops_sustained": 1400e12, "source": "fallback"}
    if p.exists():
        try:
            d = json.loads(p.read_text())
            out = {"hbm_bytes_per_s": float(d["hbm_gbs"]) * 1e9, "bf16_flops": float(d["bf16_tflops"]) * 1e12,
                   "bf16_flops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])) * 1e12,
                   "source": "measured"}
        except Exception:  # noqa: BLE001
            pass
    return out

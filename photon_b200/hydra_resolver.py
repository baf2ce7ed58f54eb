"""One-shot config resolver CLI — same contract as the reference
(ref: photon/hydra_resolver.py:30-39): ``python -m photon_b200.hydra_resolver
<overrides>`` composes ``conf/base.yaml`` with its defaults and the CLI
overrides, resolves interpolations, validates against the schema and writes
``$PHOTON_SAVE_PATH/config.yaml``. Every other process only *loads* that file.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

from photon_b200.config import compose, save_yaml


def main(argv: list[str] | None = None) -> Path:
    argv = list(sys.argv[1:] if argv is None else argv)
    save_path = os.environ.get("PHOTON_SAVE_PATH")
    if not save_path:
        raise SystemExit("PHOTON_SAVE_PATH must be set (ref: photon/hydra_resolver.py:33)")
    # Hydra's own flags: ``--config-dir/--config-path/-cd/-cp DIR`` (e.g. a reference checkout's photon/conf) and
    # ``--config-name/-cn NAME``; ``PHOTON_CONFIG_DIR`` does the same for launch scripts
    config_dir: str | None = os.environ.get("PHOTON_CONFIG_DIR") or None
    config_name = "base"
    rest: list[str] = []
    it = iter(argv)
    for a in it:
        key, _, val = a.partition("=")
        if key in ("--config-dir", "--config-path", "-cd", "-cp"):
            config_dir = val or next(it)
        elif key in ("--config-name", "-cn"):
            config_name = (val or next(it)).removesuffix(".yaml")
        else:
            rest.append(a)
    cfg = compose(rest, config_name=config_name, config_dir=config_dir)
    out = Path(save_path) / "config.yaml"
    save_yaml(cfg, out)
    print(f"[hydra_resolver] wrote {out}")
    return out


if __name__ == "__main__":
    main()

#!/usr/bin/env bash
# Stage the in-context-learning evaluation data where the gauntlet looks for it   (ref: scripts/prepare_eval_dataset.sh)
#
# The reference clones llm-foundry and copies its scripts/eval/local_data tree. There is no network here, so this script takes
# a directory that already holds that tree (a checkout, an unpacked archive, a mounted share) and links it into place:
#   SRC=/path/to/llm-foundry/scripts/eval/local_data bash scripts/prepare_eval_dataset.sh
# and then reports which task files of the chosen task list are present.
set -euo pipefail
PROJECT_PATH=${PROJECT_PATH:-$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)}
DST=${ICL_ROOT:-"$PROJECT_PATH/eval/local_data"}
mkdir -p "$(dirname "$DST")"
if [ -n "${SRC:-}" ]; then
  [ -d "$SRC" ] || { echo "prepare_eval_dataset.sh: SRC=$SRC is not a directory" >&2; exit 1; }
  rm -rf "$DST"; ln -s "$(cd "$SRC" && pwd)" "$DST"
  echo "prepare_eval_dataset.sh: $DST -> $SRC"
fi
cd "$PROJECT_PATH"
python - "$DST" "${ICL_TASKS:-tasks_v0.3}" <<'PY'
import sys
from pathlib import Path
from photon_b200.config.composer import DEFAULT_CONFIG_DIR, load_yaml
root, name = Path(sys.argv[1]), sys.argv[2]
tasks = (load_yaml(DEFAULT_CONFIG_DIR / "icl_tasks_config" / f"{name}.yaml") or {}).get("icl_tasks") or []
have = 0
for t in tasks:
    uri = str(t.get("dataset_uri", ""))
    rel = uri.split("local_data/")[-1] if "local_data/" in uri else uri
    ok = (root / rel).exists()
    have += ok
    print(f"  [{'ok' if ok else 'missing'}] {t.get('label')}: {root / rel}")
print(f"prepare_eval_dataset.sh: {have}/{len(tasks)} task files of {name} present under {root}")
PY

#!/usr/bin/env bash
# Print the fully-resolved config for a model size + overrides without launching anything
# (ref: scripts/set_llm_config.sh). usage: set_llm_config.sh mpt-1b fl.n_rounds=3 ...
source "$(dirname "${BASH_SOURCE[0]}")/_common.sh"
resolve "llm_config=${1:-mpt-125m}" "${@:2}" && cat "$PHOTON_SAVE_PATH/config.yaml"

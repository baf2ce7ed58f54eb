#!/usr/bin/env bash
# Run ONLY the evaluation gauntlet (ICL tasks + category averages) on a stored model   (ref: scripts/eval_gauntlet_only.sh)
# usage: PRETRAINED=model.npz [ICL_TASKS=tasks_v0.3] [GAUNTLET=eval_gauntlet_v0.3] [ICL_ROOT=eval/local_data] eval_gauntlet_only.sh [125M|1B|3B|7B]
# Task files are read from $ICL_ROOT (see prepare_eval_dataset.sh); tasks whose file is missing are reported as skipped.
export EXTERNAL_CONFIGS="centralized.eval_only=true centralized.stream_id=null pretrained_model_path=${PRETRAINED:?set PRETRAINED=model.npz} \
icl_tasks_config=${ICL_TASKS:-tasks_v0.3} eval_gauntlet_config=${GAUNTLET:-eval_gauntlet_v0.3} \
++icl_tasks_config.root_dir=${ICL_ROOT:-eval/local_data} llm_config.eval_first=true llm_config.eval_subset_num_batches=${EVAL_BATCHES:-1} ${EXTERNAL_CONFIGS:-}"
exec bash "$(dirname "${BASH_SOURCE[0]}")/centralised_training.sh" "${1:-125M}"

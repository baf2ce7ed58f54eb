#!/usr/bin/env bash
# Evaluate a stored model only (centralized.eval_only; ref: scripts/eval_125m_example.sh, eval_gauntlet_only.sh).
# usage: PRETRAINED=path/to/model.npz eval_125m_example.sh
export EXTERNAL_CONFIGS="centralized.eval_only=true pretrained_model_path=${PRETRAINED:?set PRETRAINED=model.npz} llm_config.eval_subset_num_batches=${EVAL_BATCHES:--1} ${EXTERNAL_CONFIGS:-}"
exec bash "$(dirname "${BASH_SOURCE[0]}")/centralised_training.sh" 125M

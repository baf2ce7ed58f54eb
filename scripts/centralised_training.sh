#!/usr/bin/env bash
# Non-federated (DDP) training: centralised_training.sh [125M|1B|3B|7B]   (ref: scripts/centralised_training.sh)
source "$(dirname "${BASH_SOURCE[0]}")/_common.sh"
SIZE=${1:-125M}
case "$SIZE" in 125M) LLM=mpt-125m ;; 1B) LLM=mpt-1b ;; 3B) LLM=mpt-3b ;; 7B) LLM=mpt-7b ;; *) echo "unknown size $SIZE"; exit 1 ;; esac
CFG="llm_config=$LLM llm_config.max_duration=${N_BATCHES:-10}ba llm_config.save_folder=$SAVE_PATH/$RUN_UUID/checkpoints ~llm_config.fsdp_config"
CFG="$CFG dataset/streams@dataset.train.streams=centralised dataset.train.root_local=$DATASET_CACHE_DIR/fed-c4 dataset.val.root_local=$DATASET_CACHE_DIR/fed-c4"
[ "$N_GPUS" -eq 0 ] && CFG="$CFG llm_config.precision=fp32 llm_config.model.attn_config.attn_impl=torch"
resolve $CFG
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True launch photon_b200.centralised_train 2>&1 | tee "$PHOTON_SAVE_PATH/centralised_train.log"

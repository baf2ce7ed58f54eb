#!/usr/bin/env bash
# Tokenise + pack + partition a corpus into per-client shards (ref: scripts/convert_c4_dataset.sh:44-50).
# usage: convert_c4_dataset.sh <source dir|file|hf name|synthetic://N> [num_clients]
source "$(dirname "${BASH_SOURCE[0]}")/_common.sh"
python -m photon_b200.dataset.convert_dataset_hf --dataset "${DATASET:-c4_en}" --splits ${SPLITS:-train_small val_xxsmall} \
  --source "${1:?source}" --out_root "$DATASET_CACHE_DIR/fed-c4" --num_clients "${2:-8}" --concat_tokens 2048 \
  --tokenizer EleutherAI/gpt-neox-20b --eos_text '<|endoftext|>'

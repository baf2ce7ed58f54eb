#!/usr/bin/env bash
# Shared launcher plumbing. Env contract follows the reference's scripts (PHOTON_SAVE_PATH, RUN_UUID,
# SAVE_PATH, EXTERNAL_CONFIGS, APPOINTED_CUDA_DEVICE, DATASET_CACHE_DIR; ref: scripts/fed_125m_example.sh:43-135).
set -euo pipefail
PROJECT_PATH=${PROJECT_PATH:-$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)}
cd "$PROJECT_PATH"
DATETIME=$(date +%Y%m%d-%H%M%S)
export RUN_UUID=${RUN_UUID:-"photon-b200-$DATETIME"}
export SAVE_PATH=${SAVE_PATH:-"$PROJECT_PATH/runs"}
export PHOTON_SAVE_PATH=${PHOTON_SAVE_PATH:-"$SAVE_PATH/$RUN_UUID"}
export DATASET_CACHE_DIR=${DATASET_CACHE_DIR:-"$PROJECT_PATH/data"}
export MASTER_ADDR=127.0.0.1
# real runs must not fall back to synthetic tokens when a dataset path is wrong (PHOTON_STRICT_DATA=0 to allow, e.g. smoke runs)
export PHOTON_STRICT_DATA=${PHOTON_STRICT_DATA:-1}
mkdir -p "$PHOTON_SAVE_PATH"
if command -v nvidia-smi >/dev/null 2>&1; then N_GPUS=${N_GPUS:-$(nvidia-smi -L | wc -l)}; else N_GPUS=${N_GPUS:-0}; fi
export APPOINTED_CUDA_DEVICE=${APPOINTED_CUDA_DEVICE:-$(seq -s, 0 $((N_GPUS > 0 ? N_GPUS - 1 : 0)))}
launch() { # launch <module> : one rank per GPU, plain python on CPU
  # PHOTON_FAULT_TOLERANT=1 (default for the federated entry points): ranks started by photon_b200.launch, which — unlike torchrun —
  # leaves the survivors alone when a rank dies; the control plane then rules the dead rank out and the federation goes on
  if [ "$N_GPUS" -gt 1 ] && [ "${PHOTON_FAULT_TOLERANT:-1}" = "1" ] && [ "$1" != "photon_b200.centralised_train" ]; then
    python -m photon_b200.launch --nproc "$N_GPUS" --master-addr 127.0.0.1 --master-port "${MASTER_PORT:-29500}" -m "$1"
  elif [ "$N_GPUS" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N_GPUS" --master-addr 127.0.0.1 --master-port "${MASTER_PORT:-29500}" -m "$1"
  else
    python -m "$1"
  fi
}
resolve() { python -m photon_b200.hydra_resolver run_uuid="$RUN_UUID" "$@" ${EXTERNAL_CONFIGS:-} 2>&1 | tee "$PHOTON_SAVE_PATH/hydra_resolver.log"; }

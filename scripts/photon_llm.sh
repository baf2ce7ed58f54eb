#!/usr/bin/env bash
# Generic federated launcher: photon_llm.sh [125M|1B|3B|7B]   (ref: scripts/photon_llm_125M.sh)
# One resolver process writes $PHOTON_SAVE_PATH/config.yaml, then the SPMD server/clients start on every GPU.
source "$(dirname "${BASH_SOURCE[0]}")/_common.sh"
SIZE=${1:-125M}
case "$SIZE" in 125M) LLM=mpt-125m ;; 1B) LLM=mpt-1b ;; 3B) LLM=mpt-3b ;; 7B) LLM=mpt-7b ;; *) echo "unknown size $SIZE"; exit 1 ;; esac
N_LOCAL_STEPS=${N_LOCAL_STEPS:-1}
PHOTON_CONFIG="llm_config=$LLM photon.refresh_period=60 photon.checkpoint=true photon.saving_path=$SAVE_PATH photon.resume_round=null"
PHOTON_CONFIG="$PHOTON_CONFIG llm_config.save_folder=$SAVE_PATH/$RUN_UUID/clients llm_config.save_overwrite=true"
PHOTON_CONFIG="$PHOTON_CONFIG fl.n_total_clients=1 fl.n_clients_per_round=1 fl.n_rounds=10"
PHOTON_CONFIG="$PHOTON_CONFIG fl.strategy_name=NESTOROV fl.strategy_kwargs.server_learning_rate=1.0 fl.strategy_kwargs.server_momentum=0.0"
PHOTON_CONFIG="$PHOTON_CONFIG llm_config.local_steps=${N_LOCAL_STEPS}ba llm_config.eval_subset_num_batches=1 llm_config.console_log_interval=100ba"
PHOTON_CONFIG="$PHOTON_CONFIG dataset.train.root_local=$DATASET_CACHE_DIR/fed-c4 dataset.val.root_local=$DATASET_CACHE_DIR/fed-c4"
if [ "$N_GPUS" -eq 0 ]; then  # CPU plumbing mode (ref README.md:82-84)
  PHOTON_CONFIG="$PHOTON_CONFIG llm_config.precision=fp32 llm_config.model.attn_config.attn_impl=torch"
else
  PHOTON_CONFIG="$PHOTON_CONFIG photon.comm_stack.shm=false photon.comm_stack.nvl=true"
fi
# TOPOLOGY=nodes: the reference's process shape — one host-side server driving N_NODES node managers, each with one
# persistent worker process per GPU of its share (no torchrun); default spmd = one process per GPU + fused round kernels
# + REMOTE_NODES=k FLEET_ADDRESS=host:port: k more machines join over gRPC — on each of them run
#   PHOTON_FLEET_TOKEN=... python -m photon_b200.node --server host:port      (scripts/photon_node.sh)
# (parameters go through the S3 bucket when S3_ENDPOINT_URL + AWS_* are set on every machine, inline in the messages otherwise;
#  N_NODES=0 = the server machine trains nothing itself)
if [ "${TOPOLOGY:-spmd}" = "nodes" ]; then
  if [ "${REMOTE_NODES:-0}" -gt 0 ]; then
    PHOTON_CONFIG="$PHOTON_CONFIG photon.fleet.n_remote_nodes=$REMOTE_NODES photon.fleet.address=${FLEET_ADDRESS:-0.0.0.0:9092}"
  fi
  resolve $PHOTON_CONFIG photon.topology=nodes photon.n_nodes="${N_NODES:-1}" photon.comm_stack.nvl=false photon.comm_stack.shm=true
  python -m photon_b200.server_app 2>&1 | tee "$PHOTON_SAVE_PATH/server.log"
else
  resolve $PHOTON_CONFIG
  launch photon_b200.server_app 2>&1 | tee "$PHOTON_SAVE_PATH/server.log"
fi

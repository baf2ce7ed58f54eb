#!/usr/bin/env bash
# Join a federation whose rank-0 process group lives on ANOTHER machine   (ref: scripts/photon_llm_125M_client_only.sh)
#
# The reference starts a Flower ClientApp that dials the SuperLink. Here every GPU process is a rank of one torch.distributed
# job, so "client only" = start this machine's ranks and point them at the first machine:
#   machine 0:  NNODES=2 NODE_RANK=0 MASTER_ADDR=<ip of machine 0> bash scripts/photon_llm_125M_client_only.sh
#   machine 1:  NNODES=2 NODE_RANK=1 MASTER_ADDR=<ip of machine 0> bash scripts/photon_llm_125M_client_only.sh
# Both machines must resolve the SAME config (same RUN_UUID / EXTERNAL_CONFIGS). The fused NVLink round kernel only spans one
# NVSwitch domain, so across machines the round transport is the NCCL all-reduce (photon.comm_stack.ray).
source "$(dirname "${BASH_SOURCE[0]}")/_common.sh"
: "${NNODES:?set NNODES}" "${NODE_RANK:?set NODE_RANK}" "${MASTER_ADDR:?set MASTER_ADDR (machine 0)}"
export MASTER_ADDR
N_LOCAL_STEPS=${N_LOCAL_STEPS:-1}
CFG="llm_config=mpt-125m photon.saving_path=$SAVE_PATH photon.resume_round=null photon.comm_stack.shm=false photon.comm_stack.ray=true"
CFG="$CFG llm_config.save_folder=$SAVE_PATH/$RUN_UUID/clients llm_config.save_overwrite=true llm_config.local_steps=${N_LOCAL_STEPS}ba"
CFG="$CFG fl.n_total_clients=$((NNODES * (N_GPUS > 0 ? N_GPUS : 1))) fl.n_clients_per_round=$((NNODES * (N_GPUS > 0 ? N_GPUS : 1))) fl.n_rounds=${N_ROUNDS:-10}"
CFG="$CFG dataset.train.root_local=$DATASET_CACHE_DIR/fed-c4 dataset.val.root_local=$DATASET_CACHE_DIR/fed-c4 llm_config.eval_subset_num_batches=1"
[ "$N_GPUS" -eq 0 ] && CFG="$CFG llm_config.precision=fp32 llm_config.model.attn_config.attn_impl=torch"
resolve $CFG
python -m torch.distributed.run --nnodes="$NNODES" --node-rank="$NODE_RANK" --nproc-per-node "$((N_GPUS > 0 ? N_GPUS : 1))" \
  --master-addr "$MASTER_ADDR" --master-port "${MASTER_PORT:-29500}" -m photon_b200.server_app 2>&1 | tee "$PHOTON_SAVE_PATH/node_${NODE_RANK}.log"

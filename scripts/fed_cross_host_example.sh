#!/usr/bin/env bash
# A federation over several machines, rehearsed on ONE: a host-side server with its gRPC fleet link and N_BOXES "machines" that join it
# (ref: the reference starts flower-superlink + one flower-supernode per machine, scripts/fed_125m_example.sh:113-135).
#   N_BOXES=2 GPUS_PER_BOX=4 bash scripts/fed_cross_host_example.sh      # two 4-GPU boxes carved out of an 8-GPU machine
# Each box joins as ONE node running the SPMD runtime inside (fused NVLink aggregation of the clients it trains, one pre-aggregated
# model per round towards the server); on a CPU-only machine every "box" is one process. On real machines run the server part here
# and `SPMD=1 bash scripts/photon_node.sh SERVER:PORT` on every other machine; set S3_ENDPOINT_URL + AWS_* everywhere to move the
# parameters through a bucket instead of through the link, and the same PHOTON_FLEET_TOKEN everywhere to authenticate the nodes.
source "$(dirname "${BASH_SOURCE[0]}")/_common.sh"
N_BOXES=${N_BOXES:-2}; PORT=${FLEET_PORT:-9092}; N_CLIENTS=${N_CLIENTS:-8}
if [ "$N_GPUS" -gt 0 ]; then GPUS_PER_BOX=${GPUS_PER_BOX:-$((N_GPUS / N_BOXES))}; else GPUS_PER_BOX=1; fi
E="llm_config=mpt-125m fl.n_total_clients=$N_CLIENTS fl.n_clients_per_round=$N_CLIENTS fl.n_rounds=${N_ROUNDS:-3} llm_config.local_steps=${LOCAL_STEPS:-8}ba"
E="$E llm_config.global_train_batch_size=32 llm_config.device_train_microbatch_size=auto fl.eval_period=null photon.checkpoint=false"
E="$E llm_config.save_folder=null ~llm_config.loggers.wandb ~llm_config.loggers.tensorboard use_wandb=false photon.saving_path=$SAVE_PATH"
E="$E dataset.train.root_local=${DATA_ROOT:-synthetic://c4} dataset.val.root_local=${DATA_ROOT:-synthetic://c4}"
if [ "$N_GPUS" -eq 0 ]; then E="$E llm_config.precision=fp32 llm_config.model.attn_config.attn_impl=torch"; fi
resolve $E photon.topology=nodes photon.n_nodes=0 photon.fleet.n_remote_nodes="$N_BOXES" photon.fleet.address="127.0.0.1:$PORT" \
        photon.comm_stack.nvl=false photon.comm_stack.shm=true
pids=()
for b in $(seq 0 $((N_BOXES - 1))); do
  if [ "$N_GPUS" -gt 0 ]; then
    devs=$(seq -s, $((b * GPUS_PER_BOX)) $(((b + 1) * GPUS_PER_BOX - 1)))
    CUDA_VISIBLE_DEVICES=$devs python -m photon_b200.launch --nproc "$GPUS_PER_BOX" --master-port $((29600 + 10 * b)) -m photon_b200.node -- \
      --server "127.0.0.1:$PORT" --spmd > "$PHOTON_SAVE_PATH/box_$b.log" 2>&1 &
  else
    python -m photon_b200.node --server "127.0.0.1:$PORT" --spmd > "$PHOTON_SAVE_PATH/box_$b.log" 2>&1 &
  fi
  pids+=($!)
done
python -m photon_b200.server_app 2>&1 | tee "$PHOTON_SAVE_PATH/server.log"
for p in "${pids[@]}"; do wait "$p" || true; done

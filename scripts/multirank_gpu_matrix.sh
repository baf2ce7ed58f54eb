set -u
BASE="fl.n_total_clients=4 fl.n_clients_per_round=4 llm_config.model.d_model=256 llm_config.model.n_heads=4 llm_config.model.n_layers=2 llm_config.max_seq_len=256 llm_config.global_train_batch_size=8 llm_config.device_train_microbatch_size=4 llm_config.device_eval_batch_size=8 llm_config.eval_subset_num_batches=2 llm_config.local_steps=2ba llm_config.log_to_console=false ~llm_config.loggers.wandb ~llm_config.loggers.tensorboard ~llm_config.callbacks dataset.train.root_local=synthetic://c4 dataset.val.root_local=synthetic://c4 photon.comm_stack.shm=false photon.comm_stack.nvl=true llm_config.save_folder=null"
run() { name=$1; port=$2; mod=$3; shift 3
  d=/tmp/mr/$name; mkdir -p $d
  PHOTON_SAVE_PATH=$d timeout 100 python -m photon_b200.hydra_resolver run_uuid=$name photon.saving_path=/tmp/mr/store $BASE "$@" > $d/resolve.log 2>&1 || { echo "$name: RESOLVE FAILED"; tail -3 $d/resolve.log; return; }
  PHOTON_SAVE_PATH=$d timeout 240 python -m photon_b200.launch --nproc 2 --master-port $port -m $mod > $d/run.log 2>&1
  rc=$?
  echo "$name: rc=$rc $(grep -E 'done|Error|error' $d/run.log | grep -v frame | tail -1 | cut -c1-200)"
}
rm -rf /tmp/mr
run g_ckpt 29921 photon_b200.server_app fl.n_rounds=2 fl.eval_period=1 photon.checkpoint=true photon.resume_round=null fl.strategy_name=fedadam "fl.strategy_kwargs={eta: 0.01, beta_1: 0.9, beta_2: 0.99, tau: 0.001}"
run g_ckpt 29923 photon_b200.server_app fl.n_rounds=3 fl.eval_period=1 photon.checkpoint=true photon.resume_round=-1 fl.strategy_name=fedadam "fl.strategy_kwargs={eta: 0.01, beta_1: 0.9, beta_2: 0.99, tau: 0.001}"
run g_fp8 29925 photon_b200.server_app fl.n_rounds=2 fl.eval_period=1 llm_config.precision=amp_fp8 photon.resume_round=null
run g_cen 29927 photon_b200.centralised_train llm_config.max_duration=4ba llm_config.eval_interval=2ba dataset/streams@dataset.train.streams=centralised

# 2-rank CPU (gloo) scenarios through the fault-tolerant launcher: checkpoints + resume, momenta, client failure, restore-run, centralised with eval, ray / s3 transports.
set -u
BASE="fl.n_total_clients=4 fl.n_clients_per_round=4 llm_config.model.d_model=32 llm_config.model.n_heads=2 llm_config.model.n_layers=1 llm_config.max_seq_len=16 llm_config.global_train_batch_size=4 llm_config.device_train_microbatch_size=2 llm_config.device_eval_batch_size=4 llm_config.eval_subset_num_batches=1 llm_config.local_steps=2ba llm_config.precision=fp32 llm_config.model.attn_config.attn_impl=torch llm_config.log_to_console=false ~llm_config.loggers.wandb ~llm_config.loggers.tensorboard ~llm_config.callbacks dataset.train.root_local=synthetic://c4 dataset.val.root_local=synthetic://c4"
run() { # name port module overrides...
  name=$1; port=$2; mod=$3; shift 3
  d=/tmp/mr/$name; mkdir -p $d
  PHOTON_SAVE_PATH=$d timeout 100 python -m photon_b200.hydra_resolver run_uuid=$name photon.saving_path=/tmp/mr/store $BASE "$@" > $d/resolve.log 2>&1 || { echo "$name: RESOLVE FAILED"; tail -3 $d/resolve.log; return; }
  PHOTON_SAVE_PATH=$d CUDA_VISIBLE_DEVICES= timeout 200 python -m photon_b200.launch --nproc 2 --master-port $port -m $mod > $d/run.log 2>&1
  rc=$?
  echo "$name: rc=$rc $(grep -E 'done|Error|error' $d/run.log | grep -v frame | tail -1 | cut -c1-160)"
}
rm -rf /tmp/mr
run a_ckpt 29901 photon_b200.server_app fl.n_rounds=2 fl.eval_period=1 photon.checkpoint=true photon.resume_round=null llm_config.save_folder=/tmp/mr/a_clients
run a_ckpt 29903 photon_b200.server_app fl.n_rounds=3 fl.eval_period=1 photon.checkpoint=true photon.resume_round=-1 llm_config.save_folder=/tmp/mr/a_clients
run b_momenta 29905 photon_b200.server_app fl.n_rounds=2 fl.eval_period=null fl.aggregate_momenta=true fl.reset_optimizer=false llm_config.optimizer.name=decoupled_adamw llm_config.save_folder=null photon.resume_round=null
run c_fault 29907 photon_b200.server_app fl.n_rounds=2 fl.eval_period=null fl.accept_failures_cnt=1 "fl.fault_injection={round: 1, cid: 2, kind: exception}" llm_config.save_folder=null photon.resume_round=null
run d_restore 29909 photon_b200.server_app fl.n_rounds=3 fl.eval_period=null photon.checkpoint=true photon.restore_run_uuid=a_ckpt photon.resume_round=-1 llm_config.save_folder=null
run e_cen 29911 photon_b200.centralised_train llm_config.max_duration=3ba llm_config.eval_interval=2ba dataset/streams@dataset.train.streams=centralised llm_config.save_folder=null
run f_ray 29913 photon_b200.server_app fl.n_rounds=2 fl.eval_period=1 photon.comm_stack.shm=false photon.comm_stack.ray=true llm_config.save_folder=null photon.resume_round=null fl.use_noise_scale_metric=true
run g_s3 29915 photon_b200.server_app fl.n_rounds=2 fl.eval_period=null photon.comm_stack.shm=false photon.comm_stack.s3=true llm_config.save_folder=null photon.resume_round=null

#!/usr/bin/env bash
# MPT-125M centralised example: batch 256, 5120 steps, DecoupledAdamW, warmup 100 + sqrt cooldown 240, no clipping,
# DDP with the fused NVLink all-reduce (ref: scripts/cen_125m_example.sh:35-100).
E="llm_config.global_train_batch_size=256 llm_config.scheduler.schedulers.lr.name=constant_with_sqrt_cooldown_with_warmup"
E="$E llm_config.scheduler.schedulers.lr.t_warmup=100ba ++llm_config.scheduler.schedulers.lr.t_cooldown=240ba llm_config.scheduler.schedulers.lr.t_max=5120ba"
E="$E ~llm_config.algorithms.gradient_clipping llm_config.optimizer={name: decoupled_adamw, lr: 6.0e-4, betas: [0.9, 0.95], eps: 1.0e-8, weight_decay: 0.0}"
E="$E llm_config.device_train_microbatch_size=auto ~llm_config.loggers.wandb ~llm_config.loggers.tensorboard"
export EXTERNAL_CONFIGS="$E ${EXTERNAL_CONFIGS:-}" N_BATCHES=${N_BATCHES:-5120}
exec bash "$(dirname "${BASH_SOURCE[0]}")/centralised_training.sh" 125M

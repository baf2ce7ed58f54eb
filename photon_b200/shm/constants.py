"""Segment-name suffixes (kept identical to ref photon/shm/constants.py:5-12 so tooling that
inspects /dev/shm keeps working): node-manager config/params, worker params / n_samples /
eval loss / metrics."""
NM_CONFIG_SHM = "_nm_cnf_shm"
NM_PARAMS_SHM = "_nm_par_shm"
W_PARAMS_SHM = "_w_par_shm"
W_N_SAMPLES_SHM = "_w_n_s_shm"
W_EVAL_LOSS_SHM = "_w_e_l_shm"
W_METRICS_SHM = "_w_mtr_shm"

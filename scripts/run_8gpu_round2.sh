R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multigpu.py tests/test_multiproc_gpu.py tests/test_zero3.py -m gpu -q 2>&1 | tail -12 > gpurun_out/tests_multigpu_final.txt; tail -5 gpurun_out/tests_multigpu_final.txt
for cfg in "PB_NVLS=1 PB_ROUND_WALK=0" "PB_NVLS=1 PB_ROUND_WALK=1" "PB_NVLS=0 PB_ROUND_WALK=0"; do
  env $cfg timeout 120 $R --master-port 29531 scripts/round_kernel_bench_mp.py 2>&1 | grep "^\[N=" | tee -a gpurun_out/round_kernel_8gpu.txt
done
# full parameter sharding: the 7B config on 2, 4 and 8 GPUs (ZeRO-3; NVLS reduce path from 4 GPUs)
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --mode ddp \
    --model mpt-7b --global-batch $((16 * n)) --steps 3 --warmup 3 --torch-arm off 2>&1 | tail -1 > gpurun_out/b_7b_zero3_${n}gpu.json
done

R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multigpu.py tests/test_multiproc_gpu.py -q 2>&1 | tail -12 > gpurun_out/tests_multigpu_final.txt; tail -5 gpurun_out/tests_multigpu_final.txt
for cfg in "PB_NVLS=1 PB_ROUND_WALK=0" "PB_NVLS=1 PB_ROUND_WALK=1" "PB_NVLS=0 PB_ROUND_WALK=0"; do
  env $cfg timeout 120 $R --master-port 29531 scripts/round_kernel_bench_mp.py 2>&1 | grep "^\[N=" | tee -a gpurun_out/round_kernel_8gpu.txt
done

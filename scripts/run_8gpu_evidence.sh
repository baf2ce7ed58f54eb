R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multigpu.py tests/test_multiproc_gpu.py -q 2>&1 | tail -25 > gpurun_out/tests_8gpu.log; tail -8 gpurun_out/tests_8gpu.log
timeout 300 $R --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/b_fed8.log 2>&1; tail -c 2600 gpurun_out/b_fed8.log
timeout 300 $R --master-port 29522 bench.py --gpus 8 --steps 6 --warmup 3 --mode ddp > gpurun_out/b_ddp8.log 2>&1; tail -c 2600 gpurun_out/b_ddp8.log
timeout 300 $R --master-port 29523 bench.py --gpus 8 --steps 6 --warmup 3 --mode fed4x2 > gpurun_out/b_fed4x2.log 2>&1; tail -c 2600 gpurun_out/b_fed4x2.log
timeout 300 $R --master-port 29524 bench.py --gpus 8 --steps 3 --warmup 3 --model mpt-1b --server fedadam --precision amp_fp8 > gpurun_out/b_1b_fp8_8.log 2>&1; tail -c 1800 gpurun_out/b_1b_fp8_8.log

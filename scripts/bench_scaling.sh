#!/usr/bin/env bash
# Headline benchmark at 1/2/4/8 GPUs + the reference-equivalent arm (device-timed, max over ranks).
cd "$(dirname "${BASH_SOURCE[0]}")/.."; mkdir -p gpurun_out
for n in ${GPUS:-1 2 4 8}; do
  if [ "$n" -eq 1 ]; then python bench.py --gpus 1 --steps "${STEPS:-4}" --warmup 3 "$@"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus "$n" --steps "${STEPS:-4}" --warmup 3 "$@"; fi | tee -a gpurun_out/scale.jsonl
done

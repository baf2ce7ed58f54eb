#!/usr/bin/env bash
# Host sanity checks before a run (GPU visibility, P2P matrix, shm size, clocks are NOT touched).
set -uo pipefail
nvidia-smi -L || echo "no GPUs visible: CPU plumbing mode (fp32, attn_impl=torch)"
nvidia-smi topo -m 2>/dev/null | head -20
df -h /dev/shm | tail -1
python - <<'PY'
import torch
print("torch", torch.__version__, "cuda", torch.cuda.is_available(), "gpus", torch.cuda.device_count())
for i in range(torch.cuda.device_count()):
    print(i, torch.cuda.get_device_name(i), [torch.cuda.can_device_access_peer(i, j) for j in range(torch.cuda.device_count()) if j != i])
PY

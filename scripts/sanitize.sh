#!/usr/bin/env bash
# compute-sanitizer passes over the kernel numerics tests at CI-sized shapes.
# usage (on a GPU box): scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
set -uo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."; mkdir -p gpurun_out
TOOL=${1:-memcheck}
compute-sanitizer --tool "$TOOL" --error-exitcode 3 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm or cross_entropy or embedding or colsum or optimizer or gemm_kk_bias" \
  2>&1 | tail -40 | tee "gpurun_out/sanitizer_$TOOL.log"

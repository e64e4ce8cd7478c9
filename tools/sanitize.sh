#!/usr/bin/env bash
# Race / memory checking of the CUDA kernels and the shm ring (the reference has no sanitizer
# configuration at all).  Run on a GPU box:
#   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
set -euo pipefail
TOOL=${1:-memcheck}
cd "$(dirname "$0")/.."
compute-sanitizer --tool "$TOOL" --error-exitcode 1 --launch-timeout 120 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "one_hop or arena or device_table or gather or negative or subgraph or random_walk"

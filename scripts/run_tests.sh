#!/usr/bin/env bash
# Build + run the test tiers (counterpart of the reference's scripts/run_cpp_ut.sh / run_python_ut.sh).
#   tier 0  native C++ tests (+ ThreadSanitizer)        no Python, no GPU
#   tier 1  python -m "not gpu"                         CPU, multi-process over localhost RPC / gloo
#   tier 2  python -m gpu                               one B200
#   tier 3  torchrun / spawn checks                     >= 2 GPUs: peer HBM sampling + gather, fused-layer numerics
#                                                       across shards, Feature / UnifiedTensor / Graph IPC per GPU
set -euo pipefail
cd "$(dirname "$0")/.."
python -m graphlearn_for_pytorch_b200.ops.build
scripts/run_cpp_ut.sh tsan
python -m pytest tests -q -m "not gpu"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -m pytest tests -q -m gpu
  if [ "$(python -c 'import torch;print(torch.cuda.device_count())')" -ge 2 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mp/p2p_check.py
    python tests/mp/feature_ipc_check.py
  fi
fi

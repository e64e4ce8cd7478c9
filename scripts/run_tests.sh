#!/usr/bin/env bash
# Build + run the test tiers (counterpart of the reference's scripts/run_cpp_ut.sh / run_python_ut.sh).
set -euo pipefail
cd "$(dirname "$0")/.."
python -m graphlearn_for_pytorch_b200.ops.build
python -m pytest tests -q -m "not gpu"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -m pytest tests -q -m gpu
  if [ "$(python -c 'import torch;print(torch.cuda.device_count())')" -ge 2 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mp/p2p_check.py
  fi
fi

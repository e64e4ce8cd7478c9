#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gather tests"; timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "gather or dropout or transposed" 2>&1 | grep -E "^E       Assert|^E       assert|^FAILED|passed|failed" | cut -c1-400 | head -12

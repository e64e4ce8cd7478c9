#!/bin/bash
# GPU pass 18 (1 GPU): gather backward as default -- full GPU suite, smoke, the bench line as the driver runs it, reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu18.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu18.log
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== kernel times (default = gather)"; timeout -k 10 300 python bench.py --kernel-times 2>&1 | grep -v Warning | tee gpurun_out/kernel_times_insitu_default.txt
echo "== bench.py (driver invocation)"; timeout -k 10 900 python bench.py 2>gpurun_out/bench18.err | grep '^{' | tail -1 | tee gpurun_out/bench_r2_final_1gpu.json | cut -c1-1800
echo "== bench.py --impl reference"; timeout -k 10 900 python bench.py --impl reference 2>gpurun_out/bench18_ref.err | grep '^{' | tail -1 | tee gpurun_out/bench_r2_final_reference_1gpu.json | cut -c1-1500
echo "== bench.py --path loader (fp32, same API level as the reference arm)"; timeout -k 10 600 python bench.py --path loader --dtype fp32 --no-arms 2>/dev/null | grep '^{' | tail -1 | tee gpurun_out/bench_r2_final_loader_fp32_1gpu.json | cut -c1-600

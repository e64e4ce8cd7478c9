#!/bin/bash
# GPU pass 3: full suite with the new relabel / hetero engine / regrow, step anatomy again, hetero benchmark arms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu3.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu3.log
for v in "" "GLT_B200_GATHER_BWD=1"; do
  echo "== sections [$v]"; env $v timeout -k 10 200 python bench.py --sections 2>&1 | tail -1
  echo "== pipelined bench [$v]"; env $v timeout -k 10 200 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['kernels_per_step'], d['e2e']['ms_per_step'])"
done
echo "== hetero engine"; timeout -k 10 400 python benchmarks/bench_hetero_rgnn.py --papers 400000 --feat-dim 1024 --hidden 512 2>&1 | tail -3
echo "== hetero loader (ours)"; timeout -k 10 400 python benchmarks/bench_hetero_rgnn.py --papers 400000 --feat-dim 1024 --hidden 512 --path loader --steps 10 2>&1 | tail -3
echo "== hetero reference"; timeout -k 10 600 python benchmarks/bench_hetero_rgnn.py --papers 400000 --feat-dim 1024 --hidden 512 --impl reference --steps 10 2>&1 | tail -3

#!/bin/bash
# GPU pass 2: hetero arena test, step anatomy (sections, unpipelined, launch list), reference with fp32 aggregation.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GLT_B200_EXPERIMENTAL=1
echo "== hetero tests"; timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "hetero" > gpurun_out/pytest_hetero.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_hetero.log
for v in "" "GLT_B200_GATHER_BWD=1"; do
  echo "== sections [$v]"; env $v timeout -k 10 200 python bench.py --sections 2>&1 | tail -1
  echo "== no-pipeline [$v]"; env $v timeout -k 10 200 python bench.py --steps 20 --warmup 5 --no-arms --no-pipeline --min-time 0.3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['kernels_per_step'])"
done
echo "== launch list (tc gemm + gather bwd)"
GLT_B200_GATHER_BWD=1 timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2.csv python bench.py --profile-steps 2 --no-pipeline > gpurun_out/launches_r2.out 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r2.csv 2 2>/dev/null | head -60
echo "== bench reference (fp32 aggregation)"; timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; echo "rc=$?"; cat gpurun_out/bench_ref2.json; tail -3 gpurun_out/bench_ref2.err

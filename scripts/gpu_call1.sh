#!/bin/bash
# GPU validation pass 1: new tcgen05 GEMM numerics, full GPU suite (incl. gather-style backward), bench pairings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GLT_B200_EXPERIMENTAL=1
echo "== tc_gemm tests"; timeout -k 10 240 python -m pytest tests/test_gpu_tc_gemm.py -x -q > gpurun_out/pytest_tc_gemm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_tc_gemm.log
echo "== full gpu suite, cuBLAS fallback for GEMMs"; GLT_B200_TC_GEMM=0 timeout -k 10 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_tc_gemm.py > gpurun_out/pytest_gpu_cublas.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu_cublas.log
echo "== engine tests with tcgen05 GEMMs"; timeout -k 10 400 python -m pytest tests/test_gpu_engine.py -q > gpurun_out/pytest_engine_tc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_engine_tc.log
echo "== bench ours (default)"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== bench ours gather-bwd"; GLT_B200_GATHER_BWD=1 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms > gpurun_out/bench_ours_gbwd.json 2> gpurun_out/bench_ours_gbwd.err; echo "rc=$?"; cat gpurun_out/bench_ours_gbwd.json; tail -5 gpurun_out/bench_ours_gbwd.err
echo "== bench ours cublas + gather-bwd"; GLT_B200_TC_GEMM=0 GLT_B200_GATHER_BWD=1 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms > gpurun_out/bench_ours_cublas.json 2> gpurun_out/bench_ours_cublas.err; echo "rc=$?"; cat gpurun_out/bench_ours_cublas.json; tail -5 gpurun_out/bench_ours_cublas.err
echo "== bench reference"; timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err

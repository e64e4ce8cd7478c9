#!/bin/bash
# GPU pass 6 (1 GPU): new-option tests, evidence pack (ncu --set full over the step + the library kernels),
# micro-benchmarks with reference arms, small trial runs of the 8-GPU scripts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== new tests"; timeout -k 10 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -x -k "deterministic or regrow or mxfp8 or hetero" > gpurun_out/pytest_gpu6.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu6.log
echo "== bench default"; timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
echo "== ncu full step"; timeout -k 10 500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/step_full python bench.py --profile-steps 1 --no-pipeline > gpurun_out/ncu_step.out 2>&1; echo "rc=$?"; ls -la gpurun_out/step_full.ncu-rep
python tools/ncu_summary.py gpurun_out/step_full.ncu-rep > gpurun_out/ncu_step_full_summary.txt 2>/dev/null; wc -l gpurun_out/ncu_step_full_summary.txt
echo "== ncu library kernels"; timeout -k 10 500 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_gather_vec|k_gather_mxfp8|k_sample_one_hop|k_negative_sample|k_subgraph|k_random_walk|k_nbr_prob|k_ell_to_coo|k_node2vec' -c 24 -f -o gpurun_out/lib_kernels python tools/run_misc_kernels.py > gpurun_out/ncu_lib.out 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_lib.out
python tools/ncu_summary.py gpurun_out/lib_kernels.ncu-rep > gpurun_out/ncu_lib_kernels_summary.txt 2>/dev/null; wc -l gpurun_out/ncu_lib_kernels_summary.txt
echo "== bench_sampler ours/ref"; timeout -k 10 300 python benchmarks/bench_sampler.py 2>/dev/null | tail -1; timeout -k 10 400 python benchmarks/bench_sampler.py --impl reference 2>/dev/null | tail -1
echo "== bench_feature ours/ref"; timeout -k 10 300 python benchmarks/bench_feature.py 2>/dev/null | tail -1; timeout -k 10 400 python benchmarks/bench_feature.py --impl reference 2>/dev/null | tail -1
echo "== seal trial"; timeout -k 10 300 python benchmarks/bench_seal_subgraph.py --nodes 5000000 --edges 100000000 --iters 10 2>&1 | tail -2
echo "== hetero device-gen trial"; timeout -k 10 300 python benchmarks/bench_hetero_rgnn.py --papers 400000 --feat-dim 1024 --hidden 512 --device-gen --steps 20 2>&1 | tail -2

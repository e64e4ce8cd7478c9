#!/bin/bash
# GPU pass 6b (1 GPU): MX GEMM test, overlap A/B, ncu evidence (summaries only: the reports are deleted on the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tc_gemm tests (incl. block-scaled MXFP8)"; timeout -k 10 300 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q > gpurun_out/pytest_tc6b.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_tc6b.log
echo "== engine tests"; timeout -k 10 400 python -m pytest tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/pytest_eng6b.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_eng6b.log
one() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'k/step', d['kernels_per_step'])"; }
echo "== bench overlap on"; timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | one
echo "== bench overlap off"; GLT_B200_OVERLAP_WGRAD=0 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-arms --min-time 0.5 2>/dev/null | one
echo "== bench_feature ref"; timeout -k 10 400 python benchmarks/bench_feature.py --impl reference 2>&1 | tail -1
echo "== ncu full step"; timeout -k 10 500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/step_full python bench.py --profile-steps 1 --no-pipeline > gpurun_out/ncu_step.out 2>&1; echo "rc=$?"
python tools/ncu_summary.py /tmp/step_full.ncu-rep > gpurun_out/ncu_step_full_summary.txt 2>/dev/null; wc -l gpurun_out/ncu_step_full_summary.txt
ncu -i /tmp/step_full.ncu-rep --page source --csv -k regex:k_sage_fused3 2>/dev/null | head -400 > gpurun_out/ncu_fused3_source_page.csv; wc -l gpurun_out/ncu_fused3_source_page.csv
echo "== ncu library kernels"; timeout -k 10 500 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_gather_vec|k_gather_mxfp8|k_sample_one_hop|k_negative_sample|k_subgraph|k_random_walk|k_nbr_prob|k_ell_to_coo' -c 16 -f -o /tmp/lib_kernels python tools/run_misc_kernels.py > gpurun_out/ncu_lib.out 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_lib.out
python tools/ncu_summary.py /tmp/lib_kernels.ncu-rep > gpurun_out/ncu_lib_kernels_summary.txt 2>/dev/null; wc -l gpurun_out/ncu_lib_kernels_summary.txt
du -sh gpurun_out

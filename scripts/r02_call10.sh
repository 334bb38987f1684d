#!/bin/bash
# Round 2, GPU call 10 (1 GPU): parity suite, A/B of the persistent windowed kernel, ncu captures of the final kernels.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c10_pytest_gpu.log 2>&1; rc=$?; el "pytest -m gpu" $rc; tail -8 gpurun_out/c10_pytest_gpu.log
timeout 500 python scripts/spmm_sweep.py --variants gather,win,winp,winpmv,fusedself > gpurun_out/c10_sweep_curlcurl.log 2>&1; el "sweep curlCurl" $?; tail -1 gpurun_out/c10_sweep_curlcurl.log
timeout 500 python scripts/spmm_sweep.py --op vecLapl --nvecs 1,16 --variants gather,win,winp,winpmv > gpurun_out/c10_sweep_veclapl.log 2>&1; el "sweep vecLapl" $?; tail -1 gpurun_out/c10_sweep_veclapl.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_spmm_winp' -s 6 -c 3 -o gpurun_out/c10_prof_winp python scripts/spmm_sweep.py --variants winpmv --nvecs 1,4 --reps 3 > gpurun_out/c10_ncu_winp.log 2>&1
el "ncu winp" $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_apply_fused' -s 6 -c 1 -o gpurun_out/c10_prof_fusedself python scripts/spmm_sweep.py --variants fusedself --nvecs 1 --reps 3 > gpurun_out/c10_ncu_fusedself.log 2>&1
el "ncu fusedself" $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_gram_mma|k_update_mma' -s 4 -c 6 -o gpurun_out/c10_prof_dense_mma python scripts/dense_bench.py --reps 1 --shapes 48x48,48x16,16x16 > gpurun_out/c10_ncu_dense.log 2>&1
el "ncu dense mma" $?

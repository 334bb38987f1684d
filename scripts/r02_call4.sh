#!/bin/bash
# Round 2, GPU call 4: windowed SpMM v3 (patterns in registers), TMA-staged Gram / update kernels.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest_gpu.log 2>&1; el "pytest -m gpu" $?; tail -12 gpurun_out/c4_pytest_gpu.log
timeout 400 python scripts/spmm_sweep.py --variants gather,win,win4 > gpurun_out/c4_sweep_curlcurl.log 2>&1; el "sweep curlCurl" $?; tail -1 gpurun_out/c4_sweep_curlcurl.log
timeout 400 python scripts/spmm_sweep.py --op vecLapl --nvecs 1,16 --variants gather,win,win4 > gpurun_out/c4_sweep_veclapl.log 2>&1; el "sweep vecLapl" $?; tail -1 gpurun_out/c4_sweep_veclapl.log
timeout 200 python scripts/dense_bench.py > gpurun_out/c4_dense.log 2>&1; el "dense bench" $?; tail -1 gpurun_out/c4_dense.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_spmm_win' -s 6 -c 3 -o gpurun_out/c4_prof_win python scripts/spmm_sweep.py --variants win --nvecs 1,4 --reps 3 > gpurun_out/c4_ncu_win.log 2>&1
el "ncu win" $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_gram_tma|k_update_tma' -s 4 -c 6 -o gpurun_out/c4_prof_dense python scripts/dense_bench.py --reps 1 --shapes 48x48,48x16,16x16 > gpurun_out/c4_ncu_dense.log 2>&1
el "ncu dense" $?

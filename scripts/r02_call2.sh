#!/bin/bash
# Round 2, GPU call 2: first run of the windowed SpMM (parity suite, A/B sweep, ncu), dense block-op timings + ncu.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/c2_pytest_gpu.log 2>&1; el "pytest -m gpu" $?; tail -5 gpurun_out/c2_pytest_gpu.log
timeout 400 python scripts/spmm_sweep.py > gpurun_out/c2_sweep_curlcurl.log 2>&1; el "sweep curlCurl" $?; tail -4 gpurun_out/c2_sweep_curlcurl.log
timeout 400 python scripts/spmm_sweep.py --op vecLapl --nvecs 1,16 > gpurun_out/c2_sweep_veclapl.log 2>&1; el "sweep vecLapl" $?; tail -4 gpurun_out/c2_sweep_veclapl.log
timeout 200 python scripts/dense_bench.py > gpurun_out/c2_dense.log 2>&1; el "dense bench" $?; tail -8 gpurun_out/c2_dense.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_spmm_win' -s 8 -c 2 -o gpurun_out/c2_prof_win python scripts/spmm_sweep.py --variants win --nvecs 1,4 --reps 3 > gpurun_out/c2_ncu_win.log 2>&1
el "ncu win" $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_gram_tiled|k_times_mat' -s 2 -c 6 -o gpurun_out/c2_prof_dense python scripts/dense_bench.py --reps 1 --shapes 48x48,48x16,16x16 > gpurun_out/c2_ncu_dense.log 2>&1
el "ncu dense" $?

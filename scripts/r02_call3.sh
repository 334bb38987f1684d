#!/bin/bash
# Round 2, GPU call 3: parity suite on the new code (windowed SpMM v2, projected eigensolve, Krylov variants), A/B sweep, ncu of nvec=1.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c3_pytest_gpu.log 2>&1; el "pytest -m gpu" $?; tail -25 gpurun_out/c3_pytest_gpu.log
timeout 400 python scripts/spmm_sweep.py > gpurun_out/c3_sweep_curlcurl.log 2>&1; el "sweep curlCurl" $?; tail -1 gpurun_out/c3_sweep_curlcurl.log
timeout 400 python scripts/spmm_sweep.py --op vecLapl --nvecs 1,16 --variants gather,win > gpurun_out/c3_sweep_veclapl.log 2>&1; el "sweep vecLapl" $?; tail -1 gpurun_out/c3_sweep_veclapl.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_spmm_win' -s 6 -c 3 -o gpurun_out/c3_prof_win python scripts/spmm_sweep.py --variants win --nvecs 1,4 --reps 3 > gpurun_out/c3_ncu_win.log 2>&1
el "ncu win" $?

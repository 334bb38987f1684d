#!/bin/bash
# Round 2, GPU call 12 (1 GPU): parity suite, paired-row windowed kernel A/B (+ncu).
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c12_pytest_gpu.log 2>&1; rc=$?; el "pytest -m gpu" $rc; tail -6 gpurun_out/c12_pytest_gpu.log
timeout 500 python scripts/spmm_sweep.py --variants gather,win,win2,win2mv,fusedself > gpurun_out/c12_sweep_curlcurl.log 2>&1; el "sweep curlCurl" $?; tail -1 gpurun_out/c12_sweep_curlcurl.log
timeout 500 python scripts/spmm_sweep.py --op vecLapl --nvecs 1,16 --variants gather,win,win2,win2mv > gpurun_out/c12_sweep_veclapl.log 2>&1; el "sweep vecLapl" $?; tail -1 gpurun_out/c12_sweep_veclapl.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_spmm_win2' -s 6 -c 2 -o gpurun_out/c12_prof_win2 python scripts/spmm_sweep.py --variants win2 --nvecs 1 --reps 3 > gpurun_out/c12_ncu_win2.log 2>&1
el "ncu win2" $?

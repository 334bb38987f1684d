#!/bin/bash
# GPU call 28 (last of the round): the whole assembly suite on the committed state + timing of config C4's assembly.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_asm.py -q > gpurun_out/c28_pytest_asm.log 2>&1
echo "pytest asm rc=$? $(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/c28_pytest_asm.log | cut -c1-400
timeout 60 python scripts/asm_bench_phc.py 192 > gpurun_out/c28_asm_phc.json 2> gpurun_out/c28_asm_phc.err
echo "asm phc rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/c28_asm_phc.json; tail -2 gpurun_out/c28_asm_phc.err

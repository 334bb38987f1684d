#!/bin/bash
# GPU call 22: device operator assembly -- parity tests, then timing + full-size parity at 256^3.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_asm.py -x -q > gpurun_out/c22_pytest_asm.log 2>&1
echo "pytest asm rc=$? $(( $(date +%s) - t0 ))s"; tail -15 gpurun_out/c22_pytest_asm.log
timeout 400 python scripts/asm_bench.py 256 > gpurun_out/c22_asm_bench.json 2> gpurun_out/c22_asm_bench.err
echo "asm_bench rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/c22_asm_bench.json; tail -5 gpurun_out/c22_asm_bench.err

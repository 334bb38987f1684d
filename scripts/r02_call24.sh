#!/bin/bash
# GPU call 24: device layout builder -- assembly tests (incl. layout equality with the host builder), timing at 256^3.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_asm.py -q > gpurun_out/c24_pytest_asm.log 2>&1
echo "pytest asm rc=$? $(( $(date +%s) - t0 ))s"; tail -12 gpurun_out/c24_pytest_asm.log
timeout 300 python scripts/asm_bench.py 256 --no-oracle > gpurun_out/c24_asm_bench.json 2> gpurun_out/c24_asm_bench.err
echo "asm_bench rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/c24_asm_bench.json; tail -3 gpurun_out/c24_asm_bench.err

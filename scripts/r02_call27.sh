#!/bin/bash
# GPU call 27: the dielectric inverse-permittivity operator assembled on the device (short: the round's last GPU seconds).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_asm.py -q -k "permittivity or errors" > gpurun_out/c27_pytest_eps.log 2>&1
echo "pytest eps rc=$? $(( $(date +%s) - t0 ))s"; tail -25 gpurun_out/c27_pytest_eps.log | cut -c1-400

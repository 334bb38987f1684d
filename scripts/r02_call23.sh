#!/bin/bash
# GPU call 23: device operator assembly -- all parity tests, timing + full-size parity at 256^3, bench with the eigensolve
# hierarchy assembled on the device.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_asm.py -x -q > gpurun_out/c23_pytest_asm.log 2>&1
echo "pytest asm rc=$? $(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/c23_pytest_asm.log
timeout 400 python scripts/asm_bench.py 256 > gpurun_out/c23_asm_bench.json 2> gpurun_out/c23_asm_bench.err
echo "asm_bench rc=$? $(( $(date +%s) - t0 ))s"; cat gpurun_out/c23_asm_bench.json; tail -3 gpurun_out/c23_asm_bench.err
timeout 600 python bench.py --assembly device --steps 50 --warmup 5 --no-sweep --cpu-seconds 2 > gpurun_out/c23_bench_device.json 2> gpurun_out/c23_bench_device.err
echo "bench --assembly device rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/c23_bench_device.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c23_bench_device.json").read().strip().splitlines()[-1])
    print("ms", round(d["ms_per_step"], 4), "parity", d["parity"]["ok"], "assembly", d["assembly"])
    e = d["eigensolve"]
    print("solve", e["value"], "iters", e["iterations"], "conv", e["converged"], "setup", e["host_setup_s"], e["assembly"][:20], "ev", [round(v, 5) for v in e["eigenvalues"]])
except Exception as ex:
    print("no bench line:", ex)
PY

#!/bin/bash
# Round 2, GPU call 6: parity suite (v2 windowed kernel, DMMA dense kernels), dense A/B, projection-tolerance sweep of the eigensolve.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c6_pytest_gpu.log 2>&1; rc=$?; el "pytest -m gpu" $rc; tail -8 gpurun_out/c6_pytest_gpu.log
timeout 200 python scripts/dense_bench.py > gpurun_out/c6_dense_mma.log 2>&1; el "dense bench mma" $?; tail -1 gpurun_out/c6_dense_mma.log
MXG_DENSE=fma timeout 200 python scripts/dense_bench.py > gpurun_out/c6_dense_fma.log 2>&1; el "dense bench fma" $?; tail -1 gpurun_out/c6_dense_fma.log
B="python bench.py --size 128 --no-cpu --no-sweep --steps 20 --solve-profile"
for v in "default" "--proj-tol-w 0.1" "--proj-tol-w 0.3" "--proj-max-iters-w 2" "--proj-max-iters-w 1"; do
  a=""; [ "$v" != "default" ] && a="$v"
  timeout 300 $B $a > gpurun_out/c6_b128.json 2> gpurun_out/c6_b128.err; el "bench128 [$v]" $?
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c6_b128.json").read().strip().splitlines()[-1]); s = d["eigensolve"]
    print("   ms", round(d["ms_per_step"], 4), "solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"],
          "proj cols/cg/reproj", s["projected_columns"], s["projection_cg_iterations"], s["reprojections_of_x"], "phase", {k: round(v, 3) for k, v in s["phase_s"].items()}, "ev0", s["eigenvalues"][0], "ev9", s["eigenvalues"][-1])
except Exception as e:
    print("   unreadable", e)
PY
done
timeout 900 python bench.py --solve-profile --no-sweep --proj-tol-w 0.1 > gpurun_out/c6_bench256.json 2> gpurun_out/c6_bench256.err; el "bench256 tolw 0.1" $?
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c6_bench256.json").read().strip().splitlines()[-1]); s = d["eigensolve"]
    print("   ms", round(d["ms_per_step"], 4), "layout_frac", round(d["roofline"]["layout_frac"], 3), "solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"],
          "proj cols/cg/reproj", s["projected_columns"], s["projection_cg_iterations"], s["reprojections_of_x"], "phase", {k: round(v, 3) for k, v in s["phase_s"].items()}, "parity", d["parity"], "cpu", d["cpu_baseline"])
except Exception as e:
    print("   unreadable", e)
PY

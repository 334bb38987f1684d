#!/bin/bash
# Round 2, GPU call 19 (1 GPU): solver-parameter sweep of the eigensolve leg at 256^3 (operators generated once, cached in /dev/shm).
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
B="python bench.py --no-cpu --no-sweep --steps 20"
for v in "default" "--block 12" "--block 14" "--sca-sweeps 1" "--sca-sweeps 3" "--vec-sweeps 1" "--vec-sweeps 3" "--proj-tol-x 0.01" "--block 20"; do
  a=""; [ "$v" != "default" ] && a="$v"
  timeout 400 $B $a > gpurun_out/c19_b.json 2> gpurun_out/c19_b.err; el "bench256 [$v]" $?
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c19_b.json").read().strip().splitlines()[-1]); s = d["eigensolve"]
    print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "maxres", "%.2e" % s["max_rel_residual"],
          "proj cols/cg/reproj", s["projected_columns"], s["projection_cg_iterations"], s["reprojections_of_x"], "vcycles", s["vcycles"], "ev9", round(s["eigenvalues"][-1], 6))
except Exception as e:
    print("   unreadable", e)
PY
done

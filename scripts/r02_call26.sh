#!/bin/bash
# Round 2, final 1-GPU check of the committed state: parity suite, smoke, default bench line.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c26_pytest_gpu.log 2>&1; el "pytest -m gpu" $?; tail -3 gpurun_out/c26_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c26_smoke.log 2>&1; el smoke $?; tail -1 gpurun_out/c26_smoke.log
timeout 900 python bench.py > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err; el "bench default" $?; tail -2 gpurun_out/c26_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c26_bench.json").read().strip().splitlines()[-1]); s = d["eigensolve"]
print("ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "layout_frac", round(d["roofline"]["layout_frac"], 3), "traffic", d["roofline"].get("traffic"), "parity", d["parity"]["ok"], "e2e", round(d["e2e"]["ms_per_step"], 3), "cpu", d["cpu_baseline"]["ms_per_apply"])
print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "proj", s["projected_columns"], s["projection_cg_iterations"], s["reprojections_of_x"], "setup", s["host_setup_s"], "ev", [round(e, 5) for e in s["eigenvalues"]])
PY

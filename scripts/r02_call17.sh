#!/bin/bash
# Round 2, GPU call 17 (1 GPU): default bench line with the final solver defaults + phase profile.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python bench.py > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; el "bench default" $?; tail -2 gpurun_out/c17_bench.err
timeout 600 python bench.py --solve-profile --no-cpu --no-sweep --steps 20 > gpurun_out/c17_bench_prof.json 2> gpurun_out/c17_bench_prof.err; el "bench solve profile" $?
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/c17_bench_ref.json 2> gpurun_out/c17_bench_ref.err; el "bench reference" $?
python - <<'PY'
import json
for f in ("c17_bench", "c17_bench_prof"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); s = d.get("eigensolve") or {}
        print(f, "ms", round(d["ms_per_step"], 4), "parity", (d.get("parity") or {}).get("ok"), "e2e", round(d["e2e"]["ms_per_step"], 3))
        print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "proj", s["projected_columns"], s["projection_cg_iterations"], s["reprojections_of_x"], "phase", s.get("phase_s"))
    except Exception as e:
        print(f, "unreadable:", e)
print(open("gpurun_out/c17_bench_ref.json").read()[:600])
PY

#!/bin/bash
# One-box validation of the committed state: the whole GPU suite, the smoke entry, the default bench line, then one
# `ncu --set full` capture of the dominant kernel of the same (short) bench command. Logs under gpurun_out/.
mkdir -p gpurun_out
t0=$(date +%s)
if [ "${1:-}" != "quick" ]; then   # `quick`: bench + ncu only (the suite ran in an earlier call on the same code)
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$? $(( $(date +%s) - t0 ))s"
fi
timeout 175 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-solve --no-sweep"
timeout 50 $B > gpurun_out/plain_short.log 2>&1 && \
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_spmm_dict -s 3 -c 1 -o gpurun_out/prof_dict_ilv $B > gpurun_out/ncu_dict_ilv.log 2>&1
echo "ncu rc=$? $(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "split", d["roofline"].get("kernel_ms"))
print("e2e", d["e2e"]); print("solve", {k: d["eigensolve"][k] for k in ("value", "iterations", "converged", "host_setup_s")})
print("block", d["block_applies"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY

#!/bin/bash
# Round 2, GPU call 18 (N GPUs): the final bench line on N ranks (e2e through the host-batch entry on every rank) + phase profile.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29513 bench.py --gpus $N --solve-profile > gpurun_out/c18_bench_n$N.json 2> gpurun_out/c18_bench_n$N.err; el "bench N=$N" $?; tail -3 gpurun_out/c18_bench_n$N.err
timeout 600 $TR --master-port 29514 bench.py --gpus $N --impl reference --steps 20 --warmup 3 > gpurun_out/c18_bench_ref_n$N.json 2> gpurun_out/c18_bench_ref_n$N.err; el "bench reference N=$N" $?
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c18_bench_n$N.json").read().strip().splitlines()[-1]); s = d.get("eigensolve") or {}
    print("ms", round(d["ms_per_step"], 4), "parity", (d.get("parity") or {}).get("ok"), "e2e", d["e2e"], "cpu", d.get("cpu_baseline"))
    print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "proj", s["projected_columns"], s["projection_cg_iterations"], s["reprojections_of_x"], "phase", s.get("phase_s"))
except Exception as e:
    print("unreadable:", e)
print(open("gpurun_out/c18_bench_ref_n$N.json").read()[:500])
PY

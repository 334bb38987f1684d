#!/bin/bash
# Round 2, GPU call 8 (N GPUs): bench at N with the widened pack role; fused vs graph A/B on the apply only.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29513 bench.py --gpus $N > gpurun_out/c8_bench_n$N.json 2> gpurun_out/c8_bench_n$N.err; el "bench N=$N" $?; tail -3 gpurun_out/c8_bench_n$N.err
MXG_HALO_FUSED=0 timeout 600 $TR --master-port 29514 bench.py --gpus $N --no-solve --no-cpu > gpurun_out/c8_bench_graph_n$N.json 2> gpurun_out/c8_bench_graph_n$N.err; el "bench graph path N=$N" $?
python - <<PY
import json
for f in ("c8_bench_n$N", "c8_bench_graph_n$N"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); s = d.get("eigensolve") or {}
        print(f, "ms", round(d["ms_per_step"], 4), "split", d["roofline"].get("kernel_ms"), "parity", (d.get("parity") or {}).get("ok"), "block", {k: round(v["ms_per_apply"], 4) for k, v in (d.get("block_applies") or {}).items()},
              "launches", d["gpu_launches"], "e2e", round(d["e2e"]["ms_per_step"], 3), "nvlink", d["roofline"].get("nvlink"))
        if s: print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "levels", s["levels"], "setup", s["host_setup_s"])
    except Exception as e:
        print(f, "unreadable:", e)
PY

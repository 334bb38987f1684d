#!/bin/bash
# Round 2, N-GPU validation: multi-rank parity, fused/graph/nccl A/B + timeline, full bench line at N.
#   gpurun --gpus N --timeout 1800 -- 'bash scripts/r02_callN.sh N'
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" -le 4 ]; then
timeout 600 $TR --master-port 29511 tests/multi_rank_check.py > gpurun_out/cN_multi_rank_n$N.log 2>&1; el "multi_rank_check N=$N" $?; grep -E "ok on|RANK|Error|error|assert" gpurun_out/cN_multi_rank_n$N.log | head -12
fi
timeout 600 $TR --master-port 29521 scripts/halo_timeline.py > gpurun_out/cN_timeline_n$N.json 2> gpurun_out/cN_timeline_n$N.err; el "timeline" $?; tail -2 gpurun_out/cN_timeline_n$N.err; cat gpurun_out/cN_timeline_n$N.json
timeout 900 $TR --master-port 29513 bench.py --gpus $N > gpurun_out/cN_bench_n$N.json 2> gpurun_out/cN_bench_n$N.err; el "bench N=$N" $?; tail -3 gpurun_out/cN_bench_n$N.err
python - <<PY
import json
for f in ("cN_bench_n$N",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); s = d.get("eigensolve") or {}
        print(f, "ms", round(d["ms_per_step"], 4), "parity", (d.get("parity") or {}).get("ok"), "block", {k: round(v["ms_per_apply"], 4) for k, v in (d.get("block_applies") or {}).items()},
              "launches", d["gpu_launches"], "e2e", round(d["e2e"]["ms_per_step"], 3), "nvlink", d["roofline"].get("nvlink"), "cpu", (d.get("cpu_baseline") or {}).get("ms_per_apply"))
        if s: print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "levels", s["levels"], "setup", s["host_setup_s"], "ev", [round(e, 6) for e in s["eigenvalues"]])
    except Exception as e:
        print(f, "unreadable:", e)
PY

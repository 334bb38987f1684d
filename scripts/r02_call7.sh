#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): multi-rank parity (fused single-launch apply, multigrid, projected eigensolve), bench at N=2,
# fused vs round-1 graph A/B.   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/r02_call7.sh'
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_rank_check.py > gpurun_out/c7_multi_rank_n$N.log 2>&1; el "multi_rank_check N=$N" $?; grep -E "ok on|RANK|Error|error|assert" gpurun_out/c7_multi_rank_n$N.log | head -20
MXG_HALO_FUSED=0 timeout 600 $TR --master-port 29512 tests/multi_rank_check.py > gpurun_out/c7_multi_rank_graph_n$N.log 2>&1; el "multi_rank_check (graph path) N=$N" $?; grep -E "RANK|Error|error|assert" gpurun_out/c7_multi_rank_graph_n$N.log | head -8
timeout 900 $TR --master-port 29513 bench.py --gpus $N > gpurun_out/c7_bench_n$N.json 2> gpurun_out/c7_bench_n$N.err; el "bench N=$N" $?; tail -3 gpurun_out/c7_bench_n$N.err
MXG_HALO_FUSED=0 timeout 600 $TR --master-port 29514 bench.py --gpus $N --no-solve --no-cpu > gpurun_out/c7_bench_graph_n$N.json 2> gpurun_out/c7_bench_graph_n$N.err; el "bench graph path N=$N" $?
MXG_HALO=nccl timeout 600 $TR --master-port 29515 bench.py --gpus $N --no-solve --no-cpu > gpurun_out/c7_bench_nccl_n$N.json 2> gpurun_out/c7_bench_nccl_n$N.err; el "bench nccl path N=$N" $?
python - <<PY
import json
for f in ("c7_bench_n$N", "c7_bench_graph_n$N", "c7_bench_nccl_n$N"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); s = d.get("eigensolve") or {}
        print(f, "ms", round(d["ms_per_step"], 4), "split", d["roofline"].get("kernel_ms"), "parity", d.get("parity"), "block", {k: round(v["ms_per_apply"], 4) for k, v in (d.get("block_applies") or {}).items()},
              "launches", d["gpu_launches"], "e2e", round(d["e2e"]["ms_per_step"], 3))
        if s: print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "ev", [round(e, 6) for e in s["eigenvalues"]])
    except Exception as e:
        print(f, "unreadable:", e)
PY

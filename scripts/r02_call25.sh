#!/bin/bash
# GPU call 25 (2 GPUs): multi-rank parity incl. the device-assembled, device-laid-out operators; the eigensolve with
# --assembly device on 2 ranks.   gpurun --gpus 2 --timeout 900 -- 'bash scripts/r02_call25.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tests/multi_rank_check.py > gpurun_out/c25_multi_rank_n$N.log 2>&1
echo "multi_rank_check N=$N rc=$? $(( $(date +%s) - t0 ))s"; grep -E "ok on|RANK|Error|error|assert" gpurun_out/c25_multi_rank_n$N.log | head -20
timeout 400 $TR --master-port 29522 bench.py --gpus $N --assembly device --steps 50 --warmup 5 --no-sweep --cpu-seconds 1 --no-assembly > gpurun_out/c25_bench_device_n$N.json 2> gpurun_out/c25_bench_device_n$N.err
echo "bench --assembly device N=$N rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/c25_bench_device_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c25_bench_device_n$N.json").read().strip().splitlines()[-1])
    e = d["eigensolve"]
    print("ms", round(d["ms_per_step"], 4), "parity", d["parity"]["ok"], "solve", round(e["value"], 3), "iters", e["iterations"], "conv", e["converged"],
          "setup", e["host_setup_s"], "ev", [round(v, 5) for v in e["eigenvalues"]])
except Exception as ex:
    print("no bench line:", ex)
PY

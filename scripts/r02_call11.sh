#!/bin/bash
# Round 2, GPU call 11 (N GPUs): multi-rank parity of the reworked single-launch apply + A/B + timeline.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_rank_check.py > gpurun_out/c11_multi_rank_n$N.log 2>&1; echo "multi_rank_check N=$N rc=$?"; grep -E "ok on|RANK|Error|error|assert" gpurun_out/c11_multi_rank_n$N.log | head -12
timeout 600 $TR --master-port 29521 scripts/halo_timeline.py > gpurun_out/c11_timeline_n$N.json 2> gpurun_out/c11_timeline_n$N.err; echo "timeline rc=$?"; tail -2 gpurun_out/c11_timeline_n$N.err; cat gpurun_out/c11_timeline_n$N.json

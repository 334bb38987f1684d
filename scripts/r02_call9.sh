#!/bin/bash
# Round 2, GPU call 9 (N GPUs): fused vs graph vs nccl on one box + %globaltimer timeline.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 scripts/halo_timeline.py > gpurun_out/c9_timeline_n$N.json 2> gpurun_out/c9_timeline_n$N.err; echo "timeline rc=$?"; tail -2 gpurun_out/c9_timeline_n$N.err; cat gpurun_out/c9_timeline_n$N.json

#!/bin/bash
# Runs on the GPU box (via gpurun): GPU test-suite, headline bench for both layouts, then the ncu
# launch list and one --set full capture of each SpMM kernel. Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --layout dict > gpurun_out/bench_dict.json 2> gpurun_out/bench_dict.err; echo "bench dict rc=$?"
python bench.py --layout sell --no-cpu > gpurun_out/bench_sell.json 2> gpurun_out/bench_sell.err; echo "bench sell rc=$?"
python bench.py --layout dict --nvec 4 --no-cpu --steps 100 > gpurun_out/bench_dict_b4.json 2> gpurun_out/bench_dict_b4.err; echo "bench b4 rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-cpu"
$B --layout dict > gpurun_out/plain_dict.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_dict.csv $B --layout dict > gpurun_out/ncu_dict.log 2>&1
echo "ncu launches rc=$?"
$B --layout dict > gpurun_out/plain_dict2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_spmm_dict -s 3 -c 2 -o gpurun_out/prof_dict $B --layout dict > gpurun_out/ncu_dict_full.log 2>&1
echo "ncu dict rc=$?"
$B --layout sell > gpurun_out/plain_sell.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_spmm_sell -s 3 -c 2 -o gpurun_out/prof_sell $B --layout sell > gpurun_out/ncu_sell_full.log 2>&1
echo "ncu sell rc=$?"
ls -la gpurun_out

#!/bin/bash
# Round 2, GPU call 1: baseline of the round-1 code + the ncu evidence round 1 lacked (dense block ops, smoother,
# default SpMM kernel).   gpurun --timeout 1500 -- 'bash scripts/r02_call1.sh'
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/c1_pytest_gpu.log 2>&1; el "pytest -m gpu" $?; tail -3 gpurun_out/c1_pytest_gpu.log
timeout 120 python tests/complex_solve_check.py > gpurun_out/c1_complex_solve.log 2>&1; el "complex eigensolve" $?; tail -2 gpurun_out/c1_complex_solve.log
timeout 300 python tests/ordered_map_check.py > gpurun_out/c1_ordered_maps.log 2>&1; el "ordered maps" $?; tail -2 gpurun_out/c1_ordered_maps.log
timeout 500 python bench.py > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err; el "bench default" $?
timeout 300 python bench.py --order soa --no-solve --no-cpu > gpurun_out/c1_bench_soa.json 2> gpurun_out/c1_bench_soa.err; el "bench --order soa" $?
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-solve --no-sweep"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_spmm_dict -s 3 -c 1 -o gpurun_out/c1_prof_ilv3 $B > gpurun_out/c1_ncu_ilv3.log 2>&1
el "ncu --set full ilv3" $?
S="python scripts/solve_profile.py --size 256 --iters 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gram_tiled|k_times_mat|k_cheb|k_spmm' --launch-skip 380 --launch-count 150 -o gpurun_out/c1_prof_solve $S > gpurun_out/c1_ncu_solve.log 2>&1
el "ncu --set full solve kernels" $?
python - <<'PY'
import json
for f in ("c1_bench_default", "c1_bench_soa"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "split", d["roofline"].get("kernel_ms"),
              "solve", (d.get("eigensolve") or {}).get("value"), "block", d.get("block_applies"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
ls -la gpurun_out | head -50

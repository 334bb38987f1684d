#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out gets its first run,
# then the default bench line, the component-major A/B, and the ncu evidence for the current default kernel.
#   gpurun --timeout 900 -- 'bash scripts/r02_first_call.sh'
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 120 python tests/complex_solve_check.py > gpurun_out/complex_solve.log 2>&1; el "complex eigensolve (first GPU run)" $?; tail -2 gpurun_out/complex_solve.log
timeout 300 python tests/ordered_map_check.py > gpurun_out/ordered_maps.log 2>&1; el "ordered maps (first GPU run)" $?; tail -2 gpurun_out/ordered_maps.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; el "pytest -m gpu" $?; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; el smoke $?
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; el "bench default" $?
timeout 300 python bench.py --order soa --no-solve --no-cpu > gpurun_out/bench_soa.json 2> gpurun_out/bench_soa.err; el "bench --order soa" $?
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-solve --no-sweep"
timeout 120 $B > gpurun_out/plain_short.log 2>&1 && \
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_default.csv $B > gpurun_out/ncu_list.log 2>&1
el "ncu launch list" $?
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_spmm_dict -s 3 -c 1 -o gpurun_out/prof_dict_default $B > gpurun_out/ncu_full.log 2>&1
el "ncu --set full" $?
python - <<'PY'
import json
for f in ("bench_default", "bench_soa"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "split", d["roofline"].get("kernel_ms"),
              "solve", (d.get("eigensolve") or {}).get("value"), "block", d.get("block_applies"))
    except Exception as e:
        print(f, "unreadable:", e)
PY

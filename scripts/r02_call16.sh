#!/bin/bash
# Round 2, GPU call 16 (1 GPU): the committed state -- parity suite, smoke, default bench line, ncu launch list + full capture of
# the bench command, C4 (phc-192, complex) and a crab-cavity apply with the N-rank parity check.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c16_pytest_gpu.log 2>&1; el "pytest -m gpu" $?; tail -4 gpurun_out/c16_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c16_smoke.log 2>&1; el smoke $?; tail -1 gpurun_out/c16_smoke.log
timeout 900 python bench.py > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; el "bench default" $?; tail -2 gpurun_out/c16_bench.err
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-solve --no-sweep"
timeout 120 $B > gpurun_out/c16_plain_short.log 2>&1 && \
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c16_launches_bench.csv $B > gpurun_out/c16_ncu_list.log 2>&1
el "ncu launch list" $?
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_spmm_win' -s 3 -c 1 -o gpurun_out/c16_prof_win_bench $B > gpurun_out/c16_ncu_full.log 2>&1
el "ncu --set full (bench command)" $?
timeout 600 python bench.py --workload phc --size 192 --no-solve --steps 100 > gpurun_out/c16_bench_phc192.json 2> gpurun_out/c16_bench_phc192.err; el "bench phc-192 (C4)" $?; tail -2 gpurun_out/c16_bench_phc192.err
timeout 600 python bench.py --workload crabcav --size 256 --no-solve --no-sweep --steps 100 > gpurun_out/c16_bench_crab256.json 2> gpurun_out/c16_bench_crab256.err; el "bench crabcav-256" $?; tail -2 gpurun_out/c16_bench_crab256.err
python - <<'PY'
import json
for f in ("c16_bench", "c16_bench_phc192", "c16_bench_crab256"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); s = d.get("eigensolve") or {}
        print(f, d["config"]["workload"], "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "layout_frac", round(d["roofline"]["layout_frac"], 3), "traffic", d["roofline"].get("traffic"),
              "parity", (d.get("parity") or {}).get("ok"), "e2e", round(d["e2e"]["ms_per_step"], 3), d["e2e"].get("result_equals_device_path"), "cpu", (d.get("cpu_baseline") or {}).get("ms_per_apply"), "layout", d["layout"])
        if s: print("   solve", round(s["value"], 3), "iters", s["iterations"], "conv", s["converged"], "divfree", s["all_divergence_free"], "setup", s["host_setup_s"], "gen", d["config"]["gen"])
    except Exception as e:
        print(f, "unreadable:", e)
PY

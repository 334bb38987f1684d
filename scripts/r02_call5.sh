#!/bin/bash
# Round 2, GPU call 5: v3 SpMM fix, new bench (projected eigensolve, cache, parity), phase profile of the solve.
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "$1 rc=$2 $(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c5_pytest_gpu.log 2>&1; rc=$?; el "pytest -m gpu" $rc; tail -8 gpurun_out/c5_pytest_gpu.log
if [ $rc -ne 0 ]; then
  timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_spmv.py -q -x -k "windowed_kernel_equals_gather_kernel and curlCurl" > gpurun_out/c5_sanitizer.log 2>&1; el "sanitizer" $?; grep -m 20 -E "Invalid|at 0x|by thread|Address" gpurun_out/c5_sanitizer.log
fi
timeout 400 python scripts/spmm_sweep.py --variants gather,win,win4 > gpurun_out/c5_sweep_curlcurl.log 2>&1; el "sweep curlCurl" $?; tail -1 gpurun_out/c5_sweep_curlcurl.log
timeout 400 python scripts/spmm_sweep.py --op vecLapl --nvecs 1,16 --variants gather,win,win4 > gpurun_out/c5_sweep_veclapl.log 2>&1; el "sweep vecLapl" $?; tail -1 gpurun_out/c5_sweep_veclapl.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/c5_bench_ref.json 2> gpurun_out/c5_bench_ref.err; el "bench reference" $?
timeout 900 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; el "bench default" $?; tail -3 gpurun_out/c5_bench.err
timeout 600 python bench.py --solve-profile --no-cpu --no-sweep --steps 20 > gpurun_out/c5_bench_prof.json 2> gpurun_out/c5_bench_prof.err; el "bench solve profile" $?
python - <<'PY'
import json
for f in ("c5_bench", "c5_bench_prof"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        s = d.get("eigensolve") or {}
        print(f, "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "layout_frac", round(d["roofline"]["layout_frac"], 3),
              "split", d["roofline"].get("kernel_ms"), "parity", d.get("parity"), "block", d.get("block_applies"))
        print("  solve", s.get("value"), "iters", s.get("iterations"), "conv", s.get("converged"), "ev", s.get("eigenvalues"),
              "divfree", s.get("all_divergence_free"), "phase", s.get("phase_s"), "proj", s.get("projected_columns"), s.get("projection_cg_iterations"), s.get("reprojections_of_x"), "setup", s.get("host_setup_s"), d["config"].get("gen"))
    except Exception as e:
        print(f, "unreadable:", e)
PY

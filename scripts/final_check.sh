#!/bin/bash
# One-box validation: new/changed GPU tests first, then the bench (default and component-interleaved dictionary kernel),
# then the rest of the GPU suite and the smoke entry. Logs under gpurun_out/.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 420 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_crabcav.py tests/test_gpu_magwave.py tests/test_gpu_mv.py -m gpu -q > gpurun_out/t1.log 2>&1
echo "t1 rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/t1.log
timeout 420 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$? $(( $(date +%s) - t0 ))s"
MXG_SPMV_ILV=auto timeout 300 python bench.py --no-cpu --no-solve > gpurun_out/bench_ilv.json 2> gpurun_out/bench_ilv.err
echo "bench ilv rc=$? $(( $(date +%s) - t0 ))s"
timeout 420 python -m pytest tests/test_gpu_solver.py tests/test_gpu_dielectric.py tests/test_gpu_multi.py tests/test_cpp_shims.py -m gpu -q > gpurun_out/t2.log 2>&1
echo "t2 rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/t2.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$? $(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
for f in ("bench_default", "bench_ilv"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "split", d["roofline"].get("kernel_ms"),
              "e2e_ms", round(d["e2e"]["ms_per_step"], 2), "solve", (d.get("eigensolve") or {}).get("seconds"),
              "iters", (d.get("eigensolve") or {}).get("iterations"), "block", d.get("block_applies"))
    except Exception as e:
        print(f, "unreadable:", e)
PY

"""A/B of the dictionary SpMM variants on ONE generated operator (env switches are read at mxg_crs_create):
gather kernels (MXG_SPMV_WIN=0) vs the windowed shared-memory kernel, both row assignments, nvec 1/4/10; every variant is
checked bit-for-bit against the first one. Usage: python scripts/spmm_sweep.py [--size 256] [--op curlCurl]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maxwell_b200 as mx
from oracle import oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--op", default="curlCurl")
ap.add_argument("--workload", default="pillbox")
ap.add_argument("--variants", default="gather,win,winmv")
ap.add_argument("--nvecs", default="1,4,10")
ap.add_argument("--reps", type=int, default=50)
args = ap.parse_args()

t = time.time()
sim = getattr(orc, args.workload)(args.size)
op = sim.op(args.op)
rowptr, col, val = op.arrays()
rg, cg = op.maps()
print("generated %s-%d %s rows=%d nnz=%d in %.1fs" % (args.workload, args.size, args.op, op.nrows, op.nnz, time.time() - t), flush=True)
ctx = mx.Context(0)
field = "psifield" if args.op == "scaLapl" else "bfield"
bmap = mx.MxMap(ctx, sim.num_global(field), rg)
cols = cg[col]
ENV = {"gather": {"MXG_SPMV_WIN": "0"}, "win": {"MXG_SPMV_WIN": "1"}, "winmv": {"MXG_SPMV_WIN": "1", "MXG_WIN_MAXVEC": "128"},
       "win_ilv1": {"MXG_SPMV_WIN": "1", "MXG_SPMV_ILV": "1"}, "fusedself": {"MXG_SPMV_WIN": "0", "MXG_FUSED_SELF": "1"}}
ref = {}
out = {}
for name in args.variants.split(","):
    for k in ("MXG_SPMV_WIN", "MXG_SPMV_ILV", "MXG_WIN_MAXVEC", "MXG_FUSED_SELF"):
        os.environ.pop(k, None)
    os.environ.update(ENV[name])
    t = time.time()
    A = mx.MxCrsMatrix.from_csr(bmap, bmap, rowptr, cols, val)
    st = A.stats()
    res = {"build_s": round(time.time() - t, 2), "device_bytes": st["device_bytes"]}
    for nv in [int(v) for v in args.nvecs.split(",")]:
        x = mx.MxMultiVector(bmap, nv, op.is_complex)
        y = mx.MxMultiVector(bmap, nv, op.is_complex)
        x.random(12345)
        for _ in range(5):
            A.apply(x, y)
        ctx.sync()
        ctx.event_record(0)
        for _ in range(args.reps):
            A.apply(x, y)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1) / args.reps
        yh = y.to_host()
        if nv not in ref:
            ref[nv] = yh
            same = None
        else:
            same = bool(np.array_equal(ref[nv], yh))
        res["nvec%d" % nv] = {"ms": round(ms, 4), "ms_per_vec": round(ms / nv, 4), "bit_exact_vs_first": same}
        del x, y
    out[name] = res
    print(name, json.dumps(res), flush=True)
    del A
    os.environ.pop("MXG_FUSED_SELF", None)
print(json.dumps(out))

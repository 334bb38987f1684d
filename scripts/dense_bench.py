"""Timing of the tall-skinny block operations at the eigensolver's shapes (no operator generation needed):
MvTransMv (Gram), MvTimesMatAddMv (update), MvAddMv, MvNorm, on n = --rows rows. Usage: python scripts/dense_bench.py"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maxwell_b200 as mx

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=20608881)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--shapes", default="48x48,48x32,48x16,32x16,16x16")
args = ap.parse_args()
ctx = mx.Context(0)
n = args.rows
m = mx.MxMap(ctx, n, np.arange(n, dtype=np.int64))
A = mx.MxMultiVector(m, 48)
X = mx.MxMultiVector(m, 48)
A.random(1)
X.random(2)
out = {}


def timeit(fn):
    for _ in range(2):
        fn()
    ctx.sync()
    ctx.event_record(0)
    for _ in range(args.reps):
        fn()
    ctx.event_record(1)
    return ctx.event_elapsed_ms(0, 1) / args.reps


for shp in args.shapes.split(","):
    k, b = [int(v) for v in shp.split("x")]
    Av = A.CloneView(list(range(k)))
    Xv = X.CloneView(list(range(b)))
    ms = timeit(lambda: Xv.MvTransMv(1.0, Av))
    byts = 8.0 * n * (k + b)
    out["gram_%s" % shp] = {"ms": round(ms, 3), "GB/s": round(byts / ms / 1e6, 1), "TFLOP/s": round(2.0 * n * k * b / ms / 1e9, 2)}
    B = np.asfortranarray(np.random.default_rng(0).uniform(-1, 1, (k, b)))
    ms = timeit(lambda: Xv.MvTimesMatAddMv(1.0, Av, B, 0.0))
    byts = 8.0 * n * (k + b)
    out["update_%s" % shp] = {"ms": round(ms, 3), "GB/s": round(byts / ms / 1e6, 1), "TFLOP/s": round(2.0 * n * k * b / ms / 1e9, 2)}
    print(shp, out["gram_%s" % shp], out["update_%s" % shp], flush=True)
Av = A.CloneView(list(range(16)))
Xv = X.CloneView(list(range(16)))
ms = timeit(lambda: Xv.MvAddMv(1.0, Av, -1.0, Xv))
out["axpby_16"] = {"ms": round(ms, 3), "GB/s": round(24.0 * n * 16 / ms / 1e6, 1)}
ms = timeit(lambda: Xv.MvNorm())
out["norm_16"] = {"ms": round(ms, 3), "GB/s": round(8.0 * n * 16 / ms / 1e6, 1)}
ms = timeit(lambda: Xv.assign(Av))
out["copy_16"] = {"ms": round(ms, 3), "GB/s": round(16.0 * n * 16 / ms / 1e6, 1)}
print(json.dumps(out))

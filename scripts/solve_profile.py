"""Per-phase timing of the eigensolve leg (verbose=2 makes MxSolver synchronise around phases)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse

import bench
import maxwell_b200 as mx

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--nev", type=int, default=10)
ap.add_argument("--block", type=int, default=16)
ap.add_argument("--tol", type=float, default=1e-8)
ap.add_argument("--workload", default="pillbox")
ap.add_argument("--iters", type=int, default=200)
args = ap.parse_args()
ctx = mx.Context(0)
_orig = mx.MxSolver.__init__


def _init(self, *a, **k):
    k["verbose"] = 2
    k["max_iters"] = args.iters
    _orig(self, *a, **k)


mx.MxSolver.__init__ = _init
out = bench.run_eigensolve(mx, ctx, args)
prof = (C.c_double * 4)()
mx.load_solver().mxs_last_profile(prof)
out["phase_s"] = {"apply_A": prof[0], "precond": prof[1], "gram": prof[2], "update": prof[3]}
out.pop("note")
print(json.dumps(out))

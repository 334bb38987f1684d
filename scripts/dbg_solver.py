import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import maxwell_b200 as mx
from oracle import oracle as orc
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import gpu_matrix

ctx = mx.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "vacuum"
if which == "vacuum":
    N = 16
    A, op, rmap, _ = gpu_matrix(mx, ctx, orc.vacuum(N), "vecLapl")
    s = mx.MxSolver(ctx, A, nev=9, block_size=24, tol=1e-9, max_iters=int(sys.argv[2]) if len(sys.argv) > 2 else 60, verbose=1)
    ev = s.solve()
    print(ev, (2 * N * np.sin(np.pi / N)) ** 2, s.iterations, s.converged)
else:
    from test_gpu_solver import _hierarchy
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    sizes = [n, n // 2, n // 4]
    t = time.time()
    sims, ops, maps, R, P = _hierarchy(mx, ctx, orc, orc.pillbox, sizes)
    print("setup s", time.time() - t)
    md = mx.MxMultiVector(maps[0], 1)
    md.from_host(sims[0].fracs("bfield"))
    prec = mx.MxGeoMultigridPrec(ctx, ops, R, P, smoother_sweeps=2)
    s = mx.MxSolver(ctx, ops[0], m_diag=md, prec=prec, nev=10, block_size=20, tol=1e-9, max_iters=100, verbose=1)
    ev = s.solve()
    print(ev, s.iterations, s.converged, s.seconds)

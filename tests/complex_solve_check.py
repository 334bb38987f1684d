"""Stand-alone body of tests/test_gpu_solver.py::test_complex_bloch_eigensolve_matches_analytic_spectrum."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import maxwell_b200 as mx
    from oracle import oracle as orc
    from conftest import gpu_matrix
    ctx = mx.Context(0)
    n, phi = 8, (0.9, -0.4, 0.25)
    sim = orc.vacuum(n, phase_shifts=phi)
    A, op, rmap, _ = gpu_matrix(mx, ctx, sim, "vecLapl")
    assert op.is_complex
    lam = [np.array([(2 * n * np.sin((2 * np.pi * m + p) / (2 * n))) ** 2 for m in range(n)]) for p in phi]
    scalar = np.sort((lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]).ravel())
    solver = mx.MxSolver(ctx, A, nev=6, block_size=10, tol=1e-9, max_iters=1000)
    ev = solver.solve()
    assert solver.converged == 6, solver.converged
    np.testing.assert_allclose(ev[:6], np.repeat(scalar[:2], 3), rtol=1e-8)
    assert solver.eigenvectors.is_complex
    res, _ = solver.check()
    assert res[:6].max() < 1e-7, res
    print("COMPLEX SOLVE OK: %d iterations, eigenvalues %s" % (solver.iterations, np.round(ev[:6], 8)))


if __name__ == "__main__":
    main()

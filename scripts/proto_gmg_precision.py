#!/usr/bin/env python
"""CPU prototype (numpy/scipy, development tool): does running the multigrid V-cycle in FP32 cost preconditioned-CG
iterations? Mirrors mxg_gmg: Chebyshev(2) smoothing on D^-1 A over [lmax/30, 1.1 lmax], re-discretised level operators,
P = trilinear interpolator, R = P^T / 8, coarse solve = Chebyshev(30). Uses the oracle only to generate matrices."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402


def cheb(A, dinv, b, x, degree, lmax, ratio, zero_start):
    lmin = lmax / ratio
    lmax = 1.1 * lmax
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    s1 = theta / delta
    rho_old = 1.0 / s1
    r = b.copy() if zero_start else b - A @ x
    w = dinv * r / theta
    x = w.copy() if zero_start else x + w
    for _ in range(1, degree):
        rho = 1.0 / (2.0 * s1 - rho_old)
        r = b - A @ x
        w = rho * rho_old * w + (2.0 * rho / delta) * (dinv * r)
        x = x + w
        rho_old = rho
    return x


class Gmg:
    def __init__(self, sizes, dtype):
        self.dtype = dtype
        sims = [orc.pillbox(n) for n in sizes]
        self.A = [s.op("vecLapl").scipy().astype(dtype).tocsr() for s in sims]
        self.dinv = []
        for A in self.A:
            d = A.diagonal()
            self.dinv.append(np.where(d != 0, 1.0 / np.where(d != 0, d, 1), 0).astype(dtype))
        self.P, self.R = [], []
        for f, c in zip(sims[:-1], sims[1:]):
            Rm = orc.interpolator(f, c, "bfield")          # coarse rows x fine cols
            Pm = Rm.transpose()
            P = Pm.scipy().astype(dtype).tocsr()
            self.P.append(P)
            self.R.append((P.T / 8.0).astype(dtype).tocsr())
        self.lmax = []
        for A, di in zip(self.A, self.dinv):
            v = np.random.default_rng(0).standard_normal(A.shape[0]).astype(dtype)
            for _ in range(30):
                v = di * (A @ v)
                v /= np.linalg.norm(v)
            self.lmax.append(float(v @ (di * (A @ v))))

    def vcycle(self, b, l=0):
        A, di = self.A[l], self.dinv[l]
        if l == len(self.A) - 1:
            return cheb(A, di, b, None, 30, self.lmax[l], 1000.0, True)
        x = cheb(A, di, b, None, 2, self.lmax[l], 30.0, True)
        r = b - A @ x
        xc = self.vcycle(self.R[l] @ r, l + 1)
        x = x + self.P[l] @ xc
        return cheb(A, di, b, x, 2, self.lmax[l], 30.0, False)

    def apply(self, b):
        return self.vcycle(b.astype(self.dtype)).astype(np.float64)


def pcg(A, b, prec, tol=1e-10, maxit=200):
    x = np.zeros_like(b)
    r = b.copy()
    z = prec(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        if np.linalg.norm(r) < tol * bn:
            return it
        z = prec(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return maxit


def main():
    sizes = [int(s) for s in (sys.argv[1:] or ["48", "24", "12", "6"])]
    A64 = orc.pillbox(sizes[0]).op("vecLapl").scipy().tocsr()
    # vecLapl is symmetric only after the mass scaling; CG here runs on the symmetrised operator as a proxy
    fa = orc.pillbox(sizes[0]).fracs("bfield")
    keep = fa > 0
    b = np.where(keep, np.random.default_rng(1).standard_normal(A64.shape[0]), 0.0)
    for dtype in (np.float64, np.float32):
        t = time.time()
        g = Gmg(sizes, dtype)
        its = pcg(A64, b, g.apply)
        print("V-cycle in %-8s: %3d PCG iterations to 1e-10 (setup+solve %.1f s)" % (np.dtype(dtype).name, its, time.time() - t))


if __name__ == "__main__":
    main()

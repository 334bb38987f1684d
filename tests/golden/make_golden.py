"""Generates tests/golden/*.npz. Run from the repo root: python tests/golden/make_golden.py

The reference ships no golden vectors for this path (its one test is
Anasazi's algebraic TestMultiVecTraits) and cannot be built here (no Trilinos), so these
fixtures are INDEPENDENT re-derivations, not reference outputs:

  vacuum_curlcurl_n6.npz   periodic vacuum curl-curl assembled with scipy Kronecker products
                           from the curl stencils of MxYeeDeyMittraCurlE.cpp:117-178 /
                           CurlB.cpp:113-169 and the GID rule of MxGrid.h:96-114 -- shares no
                           code with oracle/.
  vacuum_spectrum.npz      analytic eigenvalues sum_i (2/h_i sin(pi m_i/N_i))^2
  pillbox_counts.npz       map sizes / nnz of the pillbox operators at N=12,20 as produced by
                           the oracle at the commit that introduced it (regression pin)
  crabcav_counts.npz       same for the crab-cavity CSG config (cell_res 8), plus the sums of the
                           face / edge / volume fractions
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def kron_vacuum_curlcurl(N, L=1.0):
    """curlE * curlB on the periodic N^3 Yee grid, rows/cols in ascending-GID order."""
    h = L / N
    I = sp.identity(N, format="csr")
    S = sp.csr_matrix(np.roll(np.eye(N), 1, axis=1))      # (S v)[i] = v[i+1], periodic
    Dp = (S - I) / h                                      # forward difference
    Dm = (I - S.T) / h                                    # backward difference

    def d3(D, axis):
        mats = [I, I, I]
        mats[axis] = D
        return sp.kron(sp.kron(mats[0], mats[1]), mats[2], format="csr")   # x slowest, z fastest

    Z = sp.csr_matrix((N ** 3, N ** 3))
    Fx, Fy, Fz = d3(Dp, 0), d3(Dp, 1), d3(Dp, 2)
    Bx, By, Bz = d3(Dm, 0), d3(Dm, 1), d3(Dm, 2)
    # B_c0 = d_c1 E_c2 - d_c2 E_c1 (forward), E_c0 = d_c1 B_c2 - d_c2 B_c1 (backward)
    curlE = sp.bmat([[Z, -Fz, Fy], [Fz, Z, -Fx], [-Fy, Fx, Z]], format="csr")
    curlB = sp.bmat([[Z, -Bz, By], [Bz, Z, -Bx], [-By, Bx, Z]], format="csr")
    A = (curlE @ curlB).tocsr()
    # block (comp-major) ordering -> GID ordering comp + 3*cell restricted to cells < N
    n = N ** 3
    cell = np.arange(n)
    cx, cy, cz = cell // (N * N), (cell // N) % N, cell % N
    gcell = (cx * (N + 1) + cy) * (N + 1) + cz
    gid = np.concatenate([c + 3 * gcell for c in range(3)])
    order = np.argsort(gid)
    A = A[order][:, order].tocsr()
    A.sort_indices()
    return A, gid[order]


def main():
    A, gids = kron_vacuum_curlcurl(6)
    np.savez_compressed(os.path.join(HERE, "vacuum_curlcurl_n6.npz"), indptr=A.indptr, indices=A.indices,
                        data=A.data, gids=gids)
    spec = {}
    for N in (8, 32):
        m = np.arange(N)
        lam1 = (2 * N * np.sin(np.pi * m / N)) ** 2
        lam = (lam1[:, None, None] + lam1[None, :, None] + lam1[None, None, :]).ravel()
        spec["n%d" % N] = np.sort(lam)[:60]
    np.savez_compressed(os.path.join(HERE, "vacuum_spectrum.npz"), **spec)
    from oracle import oracle as orc
    counts = {}
    for N in (12, 20):
        sim = orc.pillbox(N)
        row = [len(sim.map(f)) for f in ("bfield", "efield", "psifield")]
        row += [sim.op(o).nnz for o in ("curlCurl", "gradDiv", "vecLapl", "scaLapl")]
        counts["n%d" % N] = np.asarray(row, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "pillbox_counts.npz"), **counts)
    sim = orc.crabcav(cell_res=8, pad=2)
    row = [len(sim.map(f)) for f in ("bfield", "efield", "psifield")]
    row += [sim.op(o).nnz for o in ("curlCurl", "gradDiv", "vecLapl", "scaLapl")]
    sums = [sim.fracs(f).sum() for f in ("bfield", "efield", "psifield")]
    np.savez_compressed(os.path.join(HERE, "crabcav_counts.npz"), counts=np.asarray(row, dtype=np.int64),
                        frac_sums=np.asarray(sums))


if __name__ == "__main__":
    main()

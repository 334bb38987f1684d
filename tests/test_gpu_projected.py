"""The eigensolve the reference performs: lowest modes of the PROJECTED operator (divergence-free fields only).
Reference: MxSolver.cpp:85-103 drives Anasazi on MxMagWaveOp::Apply, whose projection P b = b + gradPsi scaLapl^-1 divB M b
(MxMagWaveOp.cpp:893-924) removes the gradient fields. Here the block solver keeps its search space inside range(P).
Pin: scipy shift-invert on the oracle's curl-curl pencil, where the gradient space sits at eigenvalue exactly 0 and is
therefore deflated by the filter lambda > 0 -- a path that shares nothing with the projection code under test."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from conftest import gpu_matrix
from test_gpu_solver import _hierarchy

pytestmark = pytest.mark.gpu


def _psi_hierarchy(mx, ctx, orc, sims):
    """scaLapl on every level + trilinear psi transfers (restriction = P^T / 8), for the V-cycle of the projection solve."""
    ops, maps = [], []
    for s in sims:
        A, _, rmap, _ = gpu_matrix(mx, ctx, s, "scaLapl")
        ops.append(A)
        maps.append(rmap)
    R, P = [], []
    for l in range(len(sims) - 1):
        p = orc.interpolator(sims[l + 1], sims[l], field="psifield")
        r = p.transpose(scale=1.0 / 8.0)
        for mat, rm, cm, out in ((p, maps[l], maps[l + 1], P), (r, maps[l + 1], maps[l], R)):
            rowptr, col, val = mat.arrays()
            _, cg = mat.maps()
            out.append(mx.MxCrsMatrix.from_csr(rm, cm, rowptr, cg[col], val))
    return ops, maps, R, P


def _maxwell_reference(sim, nev, sigma):
    """nev lowest NON-ZERO eigenvalues of curlCurl b = k^2 dmA b (gradient fields are the exact null space)."""
    fa = sim.fracs("bfield")
    keep = np.where(fa > 0)[0]
    A = sim.op("curlCurl").scipy()[keep][:, keep].tocsc()
    M = sp.diags(fa[keep]).tocsc()
    ev = sla.eigsh(A, k=3 * nev, M=M, sigma=sigma, which="LM", tol=1e-13, return_eigenvectors=False)
    ev = np.sort(ev)
    ev = ev[ev > 1e-6 * ev.max()]
    assert len(ev) >= nev and ev[0] > 1.0
    return ev[:nev]


@pytest.mark.parametrize("scalar_prec", ["gmg", "jacobi"])
def test_projected_solve_returns_the_maxwell_modes(mx, ctx, orc, scalar_prec):
    n = 32 if scalar_prec == "gmg" else 24
    sizes = [n, n // 2, n // 4]
    sims, ops, maps, R, P = _hierarchy(mx, ctx, orc, orc.pillbox, sizes)
    sim = sims[0]
    fa = sim.fracs("bfield")
    md = mx.MxMultiVector(maps[0], 1)
    md.from_host(fa)
    prec = mx.MxGeoMultigridPrec(ctx, ops, R, P, smoother_sweeps=2)
    D, _, pmap, _ = gpu_matrix(mx, ctx, sim, "divB")
    G, _, _, _ = gpu_matrix(mx, ctx, sim, "gradPsi")
    sprec = None
    if scalar_prec == "gmg":
        sops, smaps, sR, sP = _psi_hierarchy(mx, ctx, orc, sims)
        S = sops[0]
        sprec = mx.MxGeoMultigridPrec(ctx, sops, sR, sP, smoother_sweeps=2, remove_const_field=True)
    else:
        S, _, _, _ = gpu_matrix(mx, ctx, sim, "scaLapl")
    CC, _, _, _ = gpu_matrix(mx, ctx, sim, "curlCurl")
    nev = 10
    s = mx.MxSolver(ctx, ops[0], m_diag=md, prec=prec, nev=nev, block_size=16, tol=1e-9, max_iters=300,
                    projection={"divB": D, "gradPsi": G, "scaLapl": S, "sca_prec": sprec})
    ev = s.solve()
    assert s.converged == nev, (s.converged, s.residuals)
    ref = _maxwell_reference(sim, nev, sigma=60.0)
    # mode count and values: the 10 lowest Maxwell modes, nothing from the grad-div spectrum (15.4, 21.2, ... at this size)
    np.testing.assert_allclose(ev, ref, rtol=1e-9)
    assert ev[0] > 30.0 and abs(ev[0] - (2.405 / 0.4) ** 2) < 1.0      # TM010
    # the reference's acceptance: residual with the curl-curl operator (MxMagWaveOp.cpp:1118-1209) and checkDivergences
    res, div = s.check(D, A=CC)
    assert np.all(res[:nev] < 1e-6), res[:nev]
    assert np.all(div[:nev] < 1e-6), div[:nev]
    assert np.all(s.violation[:nev] < 1e-6)
    assert s.iterations < 80, s.iterations


def test_projection_is_the_m_orthogonal_projector(mx, ctx, orc):
    sim = orc.pillbox(20)
    fa = sim.fracs("bfield")
    D, opD, pmap, bmap = gpu_matrix(mx, ctx, sim, "divB")
    G, opG, _, _ = gpu_matrix(mx, ctx, sim, "gradPsi")
    S, opS, _, _ = gpu_matrix(mx, ctx, sim, "scaLapl")
    md = mx.MxMultiVector(bmap, 1)
    md.from_host(fa)
    X = mx.MxMultiVector(bmap, 3)
    X.random(5)
    X.zero_unused(md)
    x0 = X.to_host()
    assert np.all(x0[fa == 0] == 0) and np.all(x0[fa > 0] != 0)
    its = mx.div_project(ctx, md, X, D, G, S, tol=1e-12)
    assert its > 0
    x1 = X.to_host()
    Dm, Gm, Sm = opD.scipy(), opG.scipy(), opS.scipy()
    M = sp.diags(fa)
    # (i) divergence free, (ii) the correction is a gradient field: CPU projection with scipy CG on the (consistent, singular:
    # constants per connected cavity region, cells without a live face) scalar system -- the null space of scaLapl is
    # invisible in gradPsi psi on the faces M sees, so the projection itself is unique there
    assert np.abs(Dm @ (M @ x1)).max() < 1e-9 * np.abs(Dm @ (M @ x0)).max()
    rhs = Dm @ (M @ x0)
    psi = np.zeros_like(rhs)
    for j in range(rhs.shape[1]):
        psi[:, j], info = sla.cg(Sm.tocsr(), rhs[:, j], rtol=1e-13, atol=0.0, maxiter=20000)
        assert info == 0
    want = x0 + Gm @ psi
    used = fa > 0
    assert np.all(x1[~used] == 0)
    # compared in the M (face-area) norm: faces with a tiny area fraction couple psi cells through tiny eigenvalues of
    # scaLapl, so the field VALUES on them are ill-determined by any residual-controlled solve (two CPU solvers differ by 5 %
    # in the plain 2-norm there) while their contribution to every physical quantity is weighted by that area
    mnorm = lambda v: np.sqrt(np.einsum("ij,i,ij->", v, fa, v))
    assert mnorm(x1 - want) < 1e-6 * mnorm(want)
    # (iii) M-orthogonality of the split
    for j in range(3):
        assert abs(x1[:, j] @ (fa * (x0[:, j] - x1[:, j]))) < 1e-9 * (x0[:, j] @ (fa * x0[:, j]))
    # idempotent
    mx.div_project(ctx, md, X, D, G, S, tol=1e-12)
    assert mnorm(X.to_host() - x1) < 1e-6 * mnorm(x1)


@pytest.mark.parametrize("lin_solver,sigma", [("cg", 0.05 * (2 * np.pi) ** 2), ("bicgstab", 45.0), ("gmres", 45.0)])
def test_magwave_apply_matches_cpu_shift_invert(mx, ctx, orc, lin_solver, sigma):
    """y = P (L - sigma M)^-1 M x on random input vs scipy splu on the oracle matrices + the CPU projection
    (MxMagWaveOp.cpp:863-929). sigma = 45 lies inside the spectrum: only the non-symmetric-capable solvers
    ("linear solver : type" gmres / bicgstab, MxMagWaveOp.cpp:326-338) apply."""
    sims, ops, maps, R, P = _hierarchy(mx, ctx, orc, orc.pillbox, [16, 8])
    sim = sims[0]
    fa = sim.fracs("bfield")
    md = mx.MxMultiVector(maps[0], 1)
    md.from_host(fa)
    prec = mx.MxGeoMultigridPrec(ctx, ops, R, P, smoother_sweeps=2)
    D, opD, pmap, _ = gpu_matrix(mx, ctx, sim, "divB")
    G, opG, _, _ = gpu_matrix(mx, ctx, sim, "gradPsi")
    S, opS, _, _ = gpu_matrix(mx, ctx, sim, "scaLapl")
    op = mx.MxMagWaveOp(ctx, ops[0], md, D, G, S, vec_prec=prec, shift=sigma, lin_tol=1e-11, lin_solver=lin_solver,
                        lin_basis=60, max_lin_iters=4000)
    X = mx.MxMultiVector(maps[0], 2)
    X.random(11)
    X.zero_unused(md)
    Y = X.Clone(2)
    op.Apply(X, Y)
    x, y = X.to_host(), Y.to_host()
    keep = np.where(fa > 0)[0]
    L = sim.op("vecLapl").scipy()[keep][:, keep].tocsc()
    M = sp.diags(fa[keep]).tocsc()
    b = np.zeros_like(x)
    b[keep] = sla.splu((L - sigma * M).tocsc()).solve(fa[keep, None] * x[keep])
    Dm, Gm, Sm = opD.scipy(), opG.scipy(), opS.scipy()
    rhs = Dm @ (fa[:, None] * b)
    psi = np.zeros_like(rhs)
    for j in range(rhs.shape[1]):
        psi[:, j], info = sla.cg(Sm.tocsr(), rhs[:, j], rtol=1e-13, atol=0.0, maxiter=20000)
        assert info == 0
    want = b + Gm @ psi
    mnorm = lambda v: np.sqrt(np.einsum("ij,i,ij->", v, fa, v))      # face-area norm (see the projector test)
    err = mnorm(y - want) / mnorm(want)
    assert err < 1e-6, (lin_solver, err, op.num_vec_lin_iters)
    assert op.num_vec_lin_iters > 0 and op.num_sca_lin_iters > 0


def test_complex_magwave_apply_bloch(mx, ctx, orc):
    """Complex instantiation (the reference class is templated on Scalar, MxMagWaveOp.cpp:825): Bloch-periodic vacuum."""
    sim = orc.vacuum(12, phase_shifts=(0.9, 0.4, 0.2))
    A, opA, bmap, _ = gpu_matrix(mx, ctx, sim, "vecLapl")
    D, opD, pmap, _ = gpu_matrix(mx, ctx, sim, "divB")
    G, opG, _, _ = gpu_matrix(mx, ctx, sim, "gradPsi")
    S, opS, _, _ = gpu_matrix(mx, ctx, sim, "scaLapl")
    md = mx.MxMultiVector(bmap, 1, True)
    md.from_host(np.ones(opA.nrows, dtype=np.complex128))
    sigma = 0.3
    op = mx.MxMagWaveOp(ctx, A, md, D, G, S, shift=sigma, lin_tol=1e-12, lin_solver="cg", max_lin_iters=3000)
    X = mx.MxMultiVector(bmap, 2, True)
    X.random(3)
    Y = X.Clone(2)
    op.Apply(X, Y)
    x, y = X.to_host(), Y.to_host()
    Lm, Dm, Gm, Sm = opA.scipy().tocsc(), opD.scipy(), opG.scipy(), opS.scipy().tocsc()
    b = sla.splu((Lm - sigma * sp.identity(Lm.shape[0], dtype=np.complex128, format="csc")).tocsc()).solve(x)
    psi = sla.splu(Sm).solve(Dm @ b)          # with a Bloch phase the scalar Laplacian is non-singular
    want = b + Gm @ psi
    assert np.linalg.norm(y - want) < 1e-8 * np.linalg.norm(want)
    assert np.abs(Dm @ y).max() < 1e-8 * np.abs(Dm @ b).max()

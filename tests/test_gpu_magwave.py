"""MxMagWaveOp::Apply on the GPU (reference src/MxMagWaveOp.cpp:825-943): y = P (L - sigma M)^-1 M x.
For a divergence-free eigenvector (A x = lambda M x) the operator returns x / (lambda - sigma); a grad-div
(curl-free) mode is removed by the projection."""
import numpy as np
import pytest

from conftest import gpu_matrix
from test_gpu_solver import _hierarchy

pytestmark = pytest.mark.gpu


def test_shift_invert_apply_on_eigenvectors(mx, ctx, orc):
    sims, ops, maps, R, P = _hierarchy(mx, ctx, orc, orc.pillbox, [24, 12, 6])
    sim = sims[0]
    fa = sim.fracs("bfield")
    md = mx.MxMultiVector(maps[0], 1)
    md.from_host(fa)
    prec = mx.MxGeoMultigridPrec(ctx, ops, R, P, smoother_sweeps=2)
    solver = mx.MxSolver(ctx, ops[0], m_diag=md, prec=prec, nev=8, block_size=14, tol=1e-10, max_iters=300)
    ev = solver.solve()
    assert solver.converged == 8
    D, _, pmap, _ = gpu_matrix(mx, ctx, sim, "divB")
    G, _, _, _ = gpu_matrix(mx, ctx, sim, "gradPsi")
    S, _, _, _ = gpu_matrix(mx, ctx, sim, "scaLapl")
    _, div = solver.check(D)
    sigma = 0.05 * (2 * np.pi) ** 2                       # the reference's automatic shift (src/mx.py:711-712), L = 1
    op = mx.MxMagWaveOp(ctx, ops[0], md, D, G, S, vec_prec=prec, shift=sigma, lin_tol=1e-11)
    X = solver.eigenvectors.CloneCopy(list(range(8)))
    Y = X.Clone(8)
    op.Apply(X, Y)
    assert op.num_vec_lin_iters > 0 and op.num_sca_lin_iters > 0
    xh, yh = X.to_host(), Y.to_host()
    mask = fa > 0
    n_div_free = 0
    for j in range(8):
        xn = np.linalg.norm(xh[mask, j])
        if div[j] < 1e-6:                                 # Maxwell mode: eigenvector of the shift-invert operator
            n_div_free += 1
            err = np.linalg.norm(yh[mask, j] - xh[mask, j] / (ev[j] - sigma)) / (xn / abs(ev[j] - sigma))
            assert err < 1e-6, (j, ev[j], err)
        else:                                             # grad-div mode: projected away
            assert np.linalg.norm(yh[mask, j]) < 1e-6 * xn / abs(ev[j] - sigma), (j, ev[j])
    assert n_div_free >= 2
    # the result of Apply is divergence free for arbitrary input: |divB M y| ~ 0
    Xr = X.Clone(2)
    Xr.random(3)
    mx.diag_mult(Xr, md, Xr.CloneCopy())                  # zero the unusable components
    Yr = X.Clone(2)
    op.Apply(Xr, Yr)
    My = Yr.Clone(2)
    mx.diag_mult(My, md, Yr)
    d = mx.MxMultiVector(pmap, 2)
    D.apply(My, d)
    assert d.norm2().max() < 1e-7 * My.norm2().max() * 24


@pytest.mark.parametrize("config", ["pillbox", "dsphmsph"])
def test_mag_to_elec(mx, ctx, orc, config):
    """MxMagWaveOp::magToElec (src/MxMagWaveOp.cpp:1237-1250): E = [invEps] curlB B, bit-identical to the oracle chain."""
    sim = orc.pillbox(16) if config == "pillbox" else orc.dsphmsph(16)
    Cb, opC, emap, bmap = gpu_matrix(mx, ctx, sim, "curlB")
    Ie = opI = None
    if config == "dsphmsph":
        Ie, opI, emap2, dmap = gpu_matrix(mx, ctx, sim, "invEps")
    mag = mx.MxMultiVector(bmap, 3)
    mag.random(17)
    elec = mx.MxMultiVector(emap, 3)
    mx.mag_to_elec(ctx, Cb, Ie, mag, elec)
    want = opC.apply(mag.to_host())
    if opI is not None:
        want = opI.apply(want)
    assert np.array_equal(elec.to_host(), want)

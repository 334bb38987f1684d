"""CPU tests of the dielectric generator (MxYeeFitInvEps restatement, oracle/mxo_ops.hpp): configs C3 (dielectric
sphere in a metal sphere, example/dsphmsph.py) and C4 (sapphire photonic crystal, example/phc-sapph-r0.37.py)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla


def _drop_zeros(A):
    A = A.tocsr().copy()
    A.eliminate_zeros()
    return A


def test_unit_permittivity_reduces_to_vacuum(orc):
    sph = orc.Shape.sphere(0.3, (0.05, 0.0, -0.02))
    s = orc.Sim(8, origin=(-0.5,) * 3, size=(1.0,) * 3, dielectrics=[(sph, np.eye(3))])
    v = orc.Sim(8, origin=(-0.5,) * 3, size=(1.0,) * 3)
    ie = s.op("invEps")
    A = ie.scipy()
    assert abs(A - sp.identity(ie.nrows)).max() < 1e-14          # identity values ...
    assert ie.nnz > ie.nrows                                      # ... with the cut rows' stencil zeros stored
    cc, cv = s.op("curlCurl"), v.op("curlCurl")
    assert cc.nnz > cv.nnz
    assert abs(_drop_zeros(cc.scipy()) - cv.scipy()).max() < 1e-10


def test_uniform_dielectric_scales_curlcurl(orc):
    big = orc.Shape.sphere(10.0, (0, 0, 0))
    s = orc.Sim(6, dielectrics=[(big, 4.0 * np.eye(3))])
    v = orc.Sim(6)
    assert abs(s.op("curlCurl").scipy() - v.op("curlCurl").scipy() / 4.0).max() < 1e-10
    gd, gv = s.op("gradDiv").scipy(), v.op("gradDiv").scipy()
    assert abs(gd - gv / 4.0).max() < 1e-10                       # invEpsVolAve = 3 / trace(eps)


def test_anisotropic_uniform_tensor(orc):
    """Everything inside a non-diagonal tensor: every row takes the 9-point path with inv(eps) entries."""
    eps = orc.SAPPHIRE
    big = orc.Shape.sphere(10.0, (0, 0, 0))
    s = orc.Sim(6, dielectrics=[(big, eps)])
    ie = s.op("invEps")
    assert np.all(np.diff(ie.arrays()[0]) == 9)
    inv = np.linalg.inv(eps)
    A = ie.scipy()
    rg, _ = ie.maps()
    comp = rg % 3
    for c in range(3):
        np.testing.assert_allclose(A.diagonal()[comp == c], inv[c, c], rtol=1e-13)
    # off-diagonal couplings average four neighbours each: row sums = sum_j inv[c, j]
    rs = np.asarray(A.sum(axis=1)).ravel()
    for c in range(3):
        np.testing.assert_allclose(rs[comp == c], inv[c].sum(), rtol=1e-12)


def test_dielectric_sphere_in_metal_sphere_modes(orc):
    """Analytic k values of example/dsphmsph.py:514-561 (eps=10, a=0.37, b=0.49): TM l=1 and TE l=1, 3-fold each."""
    ref = {"TM1": 2.7914257502896397024 ** 2, "TE1": 3.0859803649032310262 ** 2}
    errs = {}
    for N in (12, 20):
        s = orc.dsphmsph(N)
        A, M = s.op("vecLapl").scipy(), s.op("mRhs").scipy()
        d = M.diagonal()
        keep = np.where(d > 0)[0]
        w = sla.eigs(A[keep][:, keep].tocsc(), k=10, M=sp.diags(d[keep]).tocsc(), sigma=8.0, tol=1e-10, return_eigenvectors=False)
        assert abs(w.imag).max() < 1e-8
        w = np.sort(w.real)
        tm = w[np.argmin(abs(w - ref["TM1"]))]
        te = w[np.argmin(abs(w - ref["TE1"]))]
        assert (abs(w - tm) < 1e-6).sum() == 3 and (abs(w - te) < 1e-6).sum() == 3      # multiplicities
        errs[N] = (abs(tm - ref["TM1"]) / ref["TM1"], abs(te - ref["TE1"]) / ref["TE1"])
    assert errs[20][0] < 0.01 and errs[20][1] < 0.01
    assert errs[20][0] < 0.8 * errs[12][0] and errs[20][1] < 0.6 * errs[12][1]


def test_sapphire_crystal_bloch_operator(orc):
    s = orc.phc_sapphire(6, phase_shifts=(0.5, 0.3, 0.0))
    cc = s.op("curlCurl")
    assert cc.is_complex and cc.nrows == 3 * 6 ** 3
    assert np.diff(cc.arrays()[0]).max() > 13                     # 9-point eps^-1 rows widen the stencil
    w = np.linalg.eigvals(cc.scipy().toarray())
    w = w[abs(w) > 1e-6]
    low = w[np.argsort(abs(w))][:20]
    assert abs(low.imag).max() < 1e-9                             # the FIT eps^-1 is not symmetric, but the low bands are real
    nz = np.sort(low.real)
    assert 0.1 < nz[0] < 0.5 and nz[1] - nz[0] < 0.05             # two acoustic bands near k = (0.5, 0.3, 0)
    s0 = orc.phc_sapphire(6)
    assert not s0.op("curlCurl").is_complex                       # zero phase shift: real operator

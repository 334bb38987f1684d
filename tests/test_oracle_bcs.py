"""PEC / PMC boundary conditions of the oracle (reference src/MxGridField.cpp:41-190 with the field specialisations
MxYeeFit{B,E}Field / MxYeePsiField): analytic cavity spectra, and the symmetry-reduced octant of config C3
(example/run.py:14-29,169-179; example/dsphmsph.py:436-493) against the full ball at the same cell size."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla


def _lam(m, n):      # 1-D eigenvalue between two like walls, unit length
    return (2 * n * np.sin(np.pi * m / (2 * n))) ** 2


def _nonzero_spectrum(sim, k=12):
    w = np.sort(np.linalg.eigvals(sim.op("curlCurl").scipy().toarray()).real)
    return w[w > 1e-6][:k]


def test_pec_box_cavity_spectrum_is_analytic(orc):
    n = 6
    sim = orc.Sim(n, lower=(orc.PEC,) * 3, upper=(orc.PEC,) * 3)
    A = sim.op("curlCurl").scipy()
    assert abs(A - A.T).max() == 0.0
    got = _nonzero_spectrum(sim, 11)
    # cavity modes need two non-zero indices; multiplicity 1 per index triple with one zero, 2 with none
    want = sorted([2 * _lam(1, n)] * 3 + [3 * _lam(1, n)] * 2 + [_lam(2, n) + _lam(1, n)] * 6)
    np.testing.assert_allclose(got, want, rtol=1e-11)
    assert len(sim.map("psifield")) == n ** 3
    dB, cE = sim.op("divB").scipy(), sim.op("curlE").scipy()
    assert abs(dB @ cE).max() < 1e-10


def test_mixed_walls_along_one_axis(orc):
    n = 6
    per, pec, pmc = orc.PERIODIC, orc.PEC, orc.PMC
    half = (2 * n * np.sin(np.pi * 0.5 / (2 * n))) ** 2
    for lo, up, lowest in ((pec, pec, _lam(1, n)), (pmc, pmc, _lam(1, n)), (pmc, pec, half), (pec, pmc, half)):
        sim = orc.Sim(n, lower=(lo, per, per), upper=(up, per, per))
        got = _nonzero_spectrum(sim, 2)
        np.testing.assert_allclose(got, [lowest, lowest], rtol=1e-10)      # two polarisations, uniform in y and z


def test_octant_lower_bc_table():
    from oracle import oracle as orc
    P, M = orc.PEC, orc.PMC
    assert orc.oct_lower_bcs("TM", 1, 0) == (M, M, P)
    assert orc.oct_lower_bcs("TE", 1, 0) == (P, P, M)
    assert orc.oct_lower_bcs("TM", 1, 1) == (P, M, M)
    assert orc.oct_lower_bcs("TM", 1, 1, "im") == (M, P, M)
    assert orc.oct_lower_bcs("TE", 2, 1, "im") == (P, M, M)


def _pencil_spectrum(sim, k, sigma):
    A, M = sim.op("vecLapl").scipy(), sim.op("mRhs").scipy()
    d = M.diagonal()
    keep = np.where(d > 0)[0]
    w = sla.eigs(A[keep][:, keep].tocsc(), k=k, M=sp.diags(d[keep]).tocsc(), sigma=sigma, tol=1e-12, return_eigenvectors=False)
    assert abs(w.imag).max() < 1e-7
    return np.sort(w.real)


def test_c3_octant_reproduces_full_ball_modes(orc):
    """The octant with the multipole's symmetry planes has the same cell size as the full ball at twice the
    resolution, and the mirror boundary rows are exact: every octant mode is a full-ball mode to solver precision."""
    full = _pencil_spectrum(orc.dsphmsph(16), 14, 8.0)
    tm1, te1 = 2.7914257502896397024 ** 2, 3.0859803649032310262 ** 2          # example/dsphmsph.py:514-561
    for pol, ref in (("TM", tm1), ("TE", te1)):
        octant = orc.dsphmsph_octant(8, orc.oct_lower_bcs(pol, 1, 0))
        w = _pencil_spectrum(octant, 6, 8.0)
        w = w[w > 1e-6]
        mode = w[np.argmin(abs(w - ref))]
        assert abs(mode - ref) / ref < 0.02
        assert abs(full - mode).min() < 1e-7 * mode, (pol, mode, full)
        for v in w[w < full.max()]:
            assert abs(full - v).min() < 1e-7 * v, (pol, v)

"""CPU tests of the oracle (test infrastructure) against analytic answers, structural
invariants the reference itself probes (src/MxMagWaveOp.cpp:644-653) and the independent
fixtures in tests/golden/."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_vacuum_maps_and_structure(orc):
    N = 8
    s = orc.vacuum(N)
    # periodic: every component with a cell index == N is excluded (MxGridField.cpp:41-77)
    assert len(s.map("bfield")) == 3 * N ** 3
    assert len(s.map("efield")) == 3 * N ** 3
    assert len(s.map("psifield")) == N ** 3
    assert s.num_global("bfield") == 3 * (N + 1) ** 3
    cc = s.op("curlCurl")
    A = cc.scipy()
    assert cc.nnz == 13 * cc.nrows
    assert np.all(np.diff(cc.arrays()[0]) == 13)
    assert abs(A - A.T).max() == 0.0
    cE, dB, gP = s.op("curlE").scipy(), s.op("divB").scipy(), s.op("gradPsi").scipy()
    assert abs(dB @ cE).max() == 0.0          # div curl = 0
    assert abs(A @ gP).max() == 0.0           # curl grad = 0
    assert s.op("gradDiv").nnz == 11 * cc.nrows
    assert s.op("vecLapl").nnz == 7 * cc.nrows   # 15 structural, 7 survive purgeZeros (MxCrsMatrix.cpp:84-117)
    assert s.op("scaLapl").nnz == 7 * N ** 3


def test_vacuum_matches_independent_kronecker_assembly(orc):
    g = np.load(os.path.join(GOLD, "vacuum_curlcurl_n6.npz"))
    s = orc.vacuum(6)
    cc = s.op("curlCurl")
    rowptr, col, val = cc.arrays()
    rg, _ = cc.maps()
    assert np.array_equal(rg, g["gids"])                 # DOF indexing bit-exact
    assert np.array_equal(rowptr, g["indptr"])           # sparsity pattern bit-exact
    assert np.array_equal(col, g["indices"])
    np.testing.assert_allclose(val, g["data"], rtol=1e-14, atol=1e-12)


def test_vacuum_spectrum_analytic(orc):
    N = 8
    A = orc.vacuum(N).op("curlCurl").scipy().toarray()
    w = np.linalg.eigvalsh(A)
    nz = w[w > 1e-6]
    lam = (2 * N * np.sin(np.pi / N)) ** 2
    np.testing.assert_allclose(nz[:12], lam, rtol=1e-11)      # 12-fold lowest curl-curl mode
    np.testing.assert_allclose(nz[12], 2 * lam, rtol=1e-11)
    assert (w <= 1e-6).sum() == N ** 3 + 2                    # gradients + constant fields
    gold = np.load(os.path.join(GOLD, "vacuum_spectrum.npz"))["n8"]
    wl = np.linalg.eigvalsh(orc.vacuum(N).op("vecLapl").scipy().toarray())
    np.testing.assert_allclose(wl[:40], np.repeat(gold, 3)[:40], rtol=1e-10, atol=1e-9)


def test_literal_reference_mode_drops_periodic_wrap(orc):
    """As written, MxYeeFitEField::getCompFactor zeroes E components on a PERIODIC upper
    boundary (MxYeeFitEField.cpp:90-98 with MxGridField.cpp:64-74), which breaks symmetry.
    The default mode keeps the wrap; the literal mode reproduces the reference."""
    N = 6
    lit = orc.vacuum(N, literal=True).op("curlCurl")
    fix = orc.vacuum(N).op("curlCurl")
    assert lit.nnz < fix.nnz
    A = lit.scipy()
    assert abs(A - A.T).max() > 0
    # rows away from the upper boundaries are identical in both modes
    rg, _ = fix.maps()
    cell = rg // 3
    n1 = N + 1
    cx, cy, cz = cell // (n1 * n1), (cell // n1) % n1, cell % n1
    inner = (cx < N - 1) & (cy < N - 1) & (cz < N - 1)
    Af = fix.scipy()
    assert abs(Af[inner] - A[inner]).max() == 0.0


def test_bloch_periodic_is_hermitian_with_shifted_spectrum(orc):
    N = 6
    phi = (0.7, -0.3, 1.1)
    s = orc.vacuum(N, phase_shifts=phi)
    cc = s.op("curlCurl")
    assert cc.is_complex
    A = cc.scipy()
    assert abs(A - A.conj().T).max() < 1e-12
    w = np.linalg.eigvalsh(A.toarray())
    m = np.arange(N)
    lam = [(2 * N * np.sin((2 * np.pi * m + p) / (2 * N))) ** 2 for p in phi]
    full = np.sort((lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]).ravel())
    nz = w[w > 1e-6]
    np.testing.assert_allclose(nz[:2], full[0], rtol=1e-10)     # two transverse polarisations per k
    # K form (the reference's storage, MxCrsMatrix.cpp:145-170) gives the same product
    K = cc.kform()
    assert K.nrows == 2 * cc.nrows and K.nnz == 4 * cc.nnz
    rng = np.random.default_rng(1)
    x = rng.standard_normal(cc.ncols) + 1j * rng.standard_normal(cc.ncols)
    yk = K.apply(x.view(np.float64)).view(np.complex128)
    np.testing.assert_allclose(cc.apply(x), yk, rtol=1e-13, atol=1e-10)


def test_pillbox_counts_and_fractions(orc):
    gold = np.load(os.path.join(GOLD, "pillbox_counts.npz"))
    for N in (12, 20):
        s = orc.pillbox(N)
        got = [len(s.map(f)) for f in ("bfield", "efield", "psifield")]
        got += [s.op(o).nnz for o in ("curlCurl", "gradDiv", "vecLapl", "scaLapl")]
        assert got == list(gold["n%d" % N])
    s = orc.pillbox(20)
    fa, fl, fv = s.fracs("bfield"), s.fracs("efield"), s.fracs("psifield")
    for f in (fa, fl, fv):
        assert f.min() >= 0.0 and f.max() <= 1.0 + 1e-12
    assert (fl > 0).all()                     # E components with zero length never enter the map
    # B keeps all three components of a cell when any is usable -> some zero areas, stored explicitly
    assert (fa == 0).sum() > 0
    M = s.op("mRhs").scipy()
    assert np.array_equal(M.diagonal(), fa)
    cE = s.op("curlE")
    empty = np.diff(cE.arrays()[0]) == 0
    assert empty[fa == 0].all()               # rows of unusable B components are skipped (CurlE.cpp:148-150)
    # total metal-free volume ~ pi R^2 L
    vol = fv.sum() * (1.0 / 20) ** 3
    assert abs(vol - np.pi * 0.16 * 0.8) / (np.pi * 0.16 * 0.8) < 0.02
    dB, gP = s.op("divB").scipy(), s.op("gradPsi").scipy()
    assert abs(dB @ cE.scipy()).max() < 1e-9
    assert abs(s.op("curlCurl").scipy() @ gP).max() < 1e-7


def test_pillbox_tm010_converges_second_order(orc):
    import scipy.sparse.linalg as sla
    errs = []
    for N in (12, 24):
        s = orc.pillbox(N)
        A, M = s.op("vecLapl").scipy(), s.op("mRhs").scipy()
        d = M.diagonal()
        keep = np.where(d > 0)[0]
        w = sla.eigs(A[keep][:, keep].tocsc(), k=6, M=sp.diags(d[keep]).tocsc(), sigma=36.0, tol=1e-10,
                     return_eigenvectors=False)
        w = np.sort(w.real)
        tm010 = w[np.argmin(abs(w - (2.405 / 0.4) ** 2))]
        errs.append(abs(tm010 - (2.404825557695773 / 0.4) ** 2))
    assert errs[1] < errs[0] / 2.5          # ~4x for second order
    assert errs[1] < 0.2


def test_cut_cell_primitives(orc):
    hs = orc.Shape.halfspace((0.0, 0.0, 0.25), (0, 0, 1))         # inside: z > 0.25
    assert hs.fraction(0, 2, (1.0,), (0.0, 0.0, 0.5)) == pytest.approx(0.75)     # edge 0..1 along z
    assert hs.fraction(0, 0, (1.0,), (0.0, 0.0, 0.5)) == 1.0                     # edge along x, inside
    assert hs.fraction(0, 0, (1.0,), (0.0, 0.0, 0.0)) == 0.0
    assert hs.fraction(1, 0, (1.0, 1.0), (0.0, 0.5, 0.5)) == pytest.approx(0.75)  # x-face spans y,z in 0..1
    assert hs.fraction(1, 2, (1.0, 1.0), (0.5, 0.5, 0.5)) == 1.0                 # z-face at z=0.5
    assert hs.fraction(2, 0, (1.0, 1.0, 1.0), (0.5, 0.5, 0.5)) == pytest.approx(0.75)
    sl = orc.Shape.halfspace((0.3, 0.0, 0.0), (1, 1, 0))          # oblique cut of a z-face
    got = sl.fraction(1, 2, (1.0, 1.0), (0.5, 0.5, 0.0))
    assert got == pytest.approx(1.0 - 0.5 * 0.3 * 0.3, rel=1e-12)
    cyl = orc.Shape.cylinder(0.4, (0, 0, 1), (0, 0, 0))
    assert cyl.func((0.4, 0.0, 3.0)) == pytest.approx(0.0, abs=1e-15)
    f = cyl.fraction(0, 0, (0.1,), (0.4, 0.0, 0.0))              # edge 0.35..0.45 along x
    assert f == pytest.approx(0.5, abs=1e-10)


def test_spmm_order_and_multivector(orc):
    s = orc.pillbox(12)
    cc = s.op("curlCurl")
    rng = np.random.default_rng(0)
    X = rng.standard_normal((cc.ncols, 3))
    Y = cc.apply(X)
    rowptr, col, val = cc.arrays()
    # sequential sum in ascending column order, starting from zero (Epetra_CrsMatrix::Apply)
    for r in (0, 17, cc.nrows - 1):
        for j in range(3):
            acc = 0.0
            for p in range(rowptr[r], rowptr[r + 1]):
                acc = acc + val[p] * X[col[p], j]
            assert acc == Y[r, j]
    np.testing.assert_allclose(Y, cc.scipy() @ X, rtol=1e-13, atol=1e-9)
    assert np.array_equal(orc.csr_apply(rowptr, col, val, X), Y)

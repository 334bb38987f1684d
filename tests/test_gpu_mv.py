"""GPU multivector block ops vs numpy, mirroring the checks of Anasazi's TestMultiVecTraits that the
reference's only ctest runs (test/AnasaziInterface.cpp:40-56): clone/view aliasing semantics, MvNorm,
MvDot, MvAddMv (incl. aliased arguments), MvTransMv, MvTimesMatAddMv, SetBlock, MvScale, MvInit, MvRandom."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

N = 20011  # odd, not a multiple of any block size


def _map(mx, ctx, n=N):
    return mx.MxMap(ctx, 3 * n + 7, np.arange(n, dtype=np.int64) * 3 + 1)


def _rand(mx, m, ncols, cplx, seed):
    v = mx.MxMultiVector(m, ncols, cplx)
    v.random(seed)
    return v


@pytest.mark.parametrize("cplx", [False, True])
def test_anasazi_interface_shape_of_reference_test(mx, ctx, cplx):
    """Same shape as the reference test: map of length 5, 2 vectors, complex."""
    m = mx.MxMap(ctx, 5, np.arange(5, dtype=np.int64))
    v = mx.MxAnasaziMV(m, 2, cplx)
    assert v.GetVecLength() == 5 and v.GetNumberVecs() == 2
    v.MvInit(0.0)
    assert np.all(v.MvNorm() == 0)
    v.MvRandom()
    assert np.all(v.MvNorm() > 0)
    c = v.Clone(3)
    assert c.GetNumberVecs() == 3 and np.all(c.MvNorm() == 0)


@pytest.mark.parametrize("cplx", [False, True])
def test_random_is_keyed_by_global_id_and_column(mx, ctx, cplx):
    m = _map(mx, ctx)
    v = _rand(mx, m, 4, cplx, 12345)
    h = v.to_host()
    for j in range(4):
        ref = mx.hash_uniform(12345, m.gids, j, 0)
        if cplx:
            ref = ref + 1j * mx.hash_uniform(12345, m.gids, j, 1)
        assert np.array_equal(h[:, j], ref)
    assert np.abs(h.real).max() < 1.0
    view = v.CloneView([2])                                  # a view keeps the parent's column identity
    view.random(12345)
    assert np.array_equal(v.to_host()[:, 2], h[:, 2])


@pytest.mark.parametrize("cplx", [False, True])
def test_norm_dot(mx, ctx, cplx):
    m = _map(mx, ctx)
    a, b = _rand(mx, m, 5, cplx, 1), _rand(mx, m, 5, cplx, 2)
    ah, bh = a.to_host(), b.to_host()
    np.testing.assert_allclose(a.norm2(), np.linalg.norm(ah, axis=0), rtol=1e-14)
    ref = np.einsum("ij,ij->j", ah.conj(), bh)
    got = b.dot(a)                                           # a^H b, with the true imaginary part
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(b.MvDot(a), ref, rtol=1e-12, atol=1e-10)
    # determinism: fixed reduction tree
    assert np.array_equal(b.dot(a), got)


@pytest.mark.parametrize("cplx", [False, True])
def test_add_mv_update_scale_with_aliasing(mx, ctx, cplx):
    m = _map(mx, ctx)
    a, b, c = _rand(mx, m, 3, cplx, 1), _rand(mx, m, 3, cplx, 2), _rand(mx, m, 3, cplx, 3)
    ah, bh = a.to_host(), b.to_host()
    al, be = (0.5 - 2j, 1.5 + 0.25j) if cplx else (0.5, -1.5)
    c.MvAddMv(al, a, be, b)
    assert rel_err(c.to_host(), al * ah + be * bh) < 1e-15
    a.MvAddMv(al, a, be, b)                                  # A aliases this
    assert rel_err(a.to_host(), al * ah + be * bh) < 1e-15
    a.from_host(ah)
    b.MvAddMv(al, a, be, b)                                  # B aliases this
    assert rel_err(b.to_host(), al * ah + be * bh) < 1e-15
    b.from_host(bh)
    b.update(al, a, be)                                      # this = a*A + s*this (MxMultiVector.cpp:205-227)
    assert rel_err(b.to_host(), al * ah + be * bh) < 1e-15
    # zero coefficient drops the operand (NaNs there must not leak)
    bad = mx.MxMultiVector(m, 3, cplx)
    bad.set(float("nan"))
    c.MvAddMv(1.0, a, 0.0, bad)
    assert np.array_equal(c.to_host(), ah)
    c.MvScale(al)
    assert rel_err(c.to_host(), al * ah) < 1e-15
    s = np.array([2.0, -1.0, 0.5]) * (1 + 1j if cplx else 1)
    c.from_host(ah)
    c.MvScale(s)
    assert rel_err(c.to_host(), ah * s[None, :]) < 1e-15
    c.MvInit(3.0 - (1j if cplx else 0))
    assert np.all(c.to_host() == 3.0 - (1j if cplx else 0))
    if cplx:
        c.from_host(ah)
        c.conj()
        assert np.array_equal(c.to_host(), ah.conj())
    n0 = a.norm2()
    a.normalize()
    np.testing.assert_allclose(a.norm2(), 1.0, rtol=1e-14)
    assert rel_err(a.to_host() * n0[None, :], ah) < 1e-15


@pytest.mark.parametrize("cplx", [False, True])
def test_clone_view_copy_setblock_semantics(mx, ctx, cplx):
    m = _map(mx, ctx)
    v = _rand(mx, m, 6, cplx, 9)
    vh = v.to_host()
    cp = v.CloneCopy()
    cp2 = v.CloneCopy([4, 1])
    assert np.array_equal(cp.to_host(), vh) and np.array_equal(cp2.to_host(), vh[:, [4, 1]])
    view = v.CloneViewNonConst([5, 0, 3])                   # non-contiguous, unordered list is legal
    assert np.array_equal(view.to_host(), vh[:, [5, 0, 3]])
    view.MvInit(7.0)                                         # writes through to the parent
    after = v.to_host()
    assert np.all(after[:, [5, 0, 3]] == 7.0) and np.array_equal(after[:, [1, 2, 4]], vh[:, [1, 2, 4]])
    assert np.array_equal(cp.to_host(), vh)                  # deep copies are unaffected
    src = _rand(mx, m, 2, cplx, 10)
    v.SetBlock(src, [4, 2])                                  # SetBlock (MxAnasaziMV.cpp:201-213)
    sh = src.to_host()
    got = v.to_host()
    assert np.array_equal(got[:, 4], sh[:, 0]) and np.array_equal(got[:, 2], sh[:, 1])
    del v                                                    # the view keeps the storage alive
    assert np.all(view.to_host()[:, 0] == 7.0)
    w = cp.Clone(2)
    w.assign(cp2)
    assert np.array_equal(w.to_host(), vh[:, [4, 1]])


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("k,b", [(1, 1), (3, 2), (8, 4), (20, 10), (36, 12), (13, 7), (48, 48), (48, 33), (70, 50)])
def test_trans_mv_and_times_mat_add_mv(mx, ctx, cplx, k, b):
    m = _map(mx, ctx)
    A, X = _rand(mx, m, k, cplx, 21), _rand(mx, m, b, cplx, 22)
    Ah, Xh = A.to_host(), X.to_host()
    alpha = (0.75 + 0.5j) if cplx else -1.25
    G = X.MvTransMv(alpha, A)                               # alpha * A^H X (MxAnasaziMV.cpp:114-197)
    ref = alpha * (Ah.conj().T @ Xh)
    assert G.shape == (k, b)
    np.testing.assert_allclose(G, ref, rtol=1e-11, atol=1e-9)
    rng = np.random.default_rng(k * 100 + b)
    B = rng.standard_normal((k, b)) + (1j * rng.standard_normal((k, b)) if cplx else 0)
    beta = (0.3 - 1j) if cplx else 2.0
    Y = _rand(mx, m, b, cplx, 23)
    Yh = Y.to_host()
    Y.MvTimesMatAddMv(alpha, A, B, beta)                    # this = alpha*A*B + beta*this (MxAnasaziMV.cpp:8-86)
    assert rel_err(Y.to_host(), alpha * (Ah @ B) + beta * Yh) < 1e-13
    Y.MvTimesMatAddMv(1.0, A, B, 0.0)
    assert rel_err(Y.to_host(), Ah @ B) < 1e-13
    if cplx or k > 48:
        with pytest.raises(mx.MxError):
            A.MvTimesMatAddMv(1.0, A, np.eye(k), 0.0)        # complex / wide blocks: the result must not share columns with A
    else:
        # real: in-place right-multiplication of a basis block (a row chunk of A is complete in shared memory before
        # those rows are written), also into a subset of A's own columns
        C = rng.standard_normal((k, k))
        A.MvTimesMatAddMv(1.0, A, C, 0.0)
        assert rel_err(A.to_host(), Ah @ C) < 1e-13
        if k >= 2:
            Ah2 = A.to_host()
            sub = A.CloneView([1, 0])
            C2 = rng.standard_normal((k, 2))
            sub.MvTimesMatAddMv(0.5, A, C2, 0.0)
            want = Ah2.copy()
            want[:, [1, 0]] = 0.5 * (Ah2 @ C2)
            assert rel_err(A.to_host(), want) < 1e-13


def test_views_feed_gram_products(mx, ctx):
    m = _map(mx, ctx)
    V = _rand(mx, m, 12, False, 5)
    Vh = V.to_host()
    left, right = V.CloneView([0, 2, 4, 6]), V.CloneView([11, 1])
    G = right.MvTransMv(1.0, left)
    np.testing.assert_allclose(G, Vh[:, [0, 2, 4, 6]].T @ Vh[:, [11, 1]], rtol=1e-11, atol=1e-9)
    tgt = V.CloneViewNonConst([8, 9])
    tgt.MvTimesMatAddMv(1.0, left, np.ones((4, 2)), 0.0)
    np.testing.assert_allclose(V.to_host()[:, 8], Vh[:, [0, 2, 4, 6]].sum(axis=1), rtol=1e-13, atol=1e-12)


def test_argument_errors_are_reported(mx, ctx):
    m1, m2 = _map(mx, ctx, 100), _map(mx, ctx, 101)
    a, b = mx.MxMultiVector(m1, 2), mx.MxMultiVector(m2, 2)
    with pytest.raises(mx.MxError):
        a.MvAddMv(1.0, a, 1.0, b)
    with pytest.raises(mx.MxError):
        a.dot(mx.MxMultiVector(m1, 3))
    with pytest.raises(mx.MxError):
        a.CloneView([2])
    with pytest.raises(mx.MxError):
        mx.MxMultiVector(m1, 0)
    with pytest.raises(mx.MxError):
        mx.MxMap(ctx, 10, np.array([3, 2], dtype=np.int64))
    empty = mx.MxMap(ctx, 10, np.array([], dtype=np.int64))     # ragged: a rank may own nothing
    e = mx.MxMultiVector(empty, 2)
    e.random(1)
    assert np.all(e.norm2() == 0)


@pytest.mark.parametrize("is_complex", [False, True])
def test_to_grid_matches_mxio_layout(mx, ctx, orc, is_complex):
    """mxg_mv_to_grid: the dense [x][y][z][comp] node-grid array MxIO::save writes (src/MxIO.cpp:166-221),
    zeros at DOFs the map masks out; bit-exact scatter by GID."""
    n = 10
    sim = orc.pillbox(n)
    gids = sim.map("bfield")
    m = mx.MxMap(ctx, sim.num_global("bfield"), gids)
    x = mx.MxMultiVector(m, 3, is_complex)
    x.random(5)
    xh = x.to_host()
    total = 3 * (n + 1) ** 3
    for col in (0, 2):
        want = np.zeros(total, dtype=xh.dtype)
        want[gids] = xh[:, col]
        got = x.to_grid(col, 0, total)
        assert np.array_equal(got, want)
        arr = got.reshape(n + 1, n + 1, n + 1, 3)
        assert np.all(arr[0] == 0)                         # the x = 0 plane lies in the metal
    lo, hi = int(gids.min()), int(gids.max()) + 1          # a sub-range holding every owned DOF works too
    assert np.array_equal(x.to_grid(1, lo, hi), _dense(gids, xh[:, 1], lo, hi))
    with pytest.raises(mx.MxError):
        x.to_grid(0, lo + 5, hi)


def _dense(gids, vals, lo, hi):
    out = np.zeros(hi - lo, dtype=vals.dtype)
    out[gids - lo] = vals
    return out


def test_axpby_cols_remove_const_zero_unused(mx, ctx):
    """Per-column fused update (the CG / Chebyshev recurrences), removeConstField (MxGeoMultigridPrec.cpp:400-411,
    MxUtil.cpp:483-503) and MxGridField::zeroUnusedComponents (MxGridField.cpp:548-576)."""
    n = 10007
    m = mx.MxMap(ctx, n, np.arange(n, dtype=np.int64))
    for cx in (False, True):
        A, B, Dst = (mx.MxMultiVector(m, 3, cx) for _ in range(3))
        A.random(1)
        B.random(2)
        a, b = A.to_host(), B.to_host()
        al = np.array([0.5, -2.0, 0.0]) + (1j * np.array([0.25, 0.0, 1.0]) if cx else 0)
        be = np.array([1.0, 3.0, -1.5]) + (1j * np.array([0.0, -0.5, 2.0]) if cx else 0)
        Dst.axpby_cols(al, A, be, B)
        np.testing.assert_allclose(Dst.to_host(), a * al[None, :] + b * be[None, :], rtol=1e-14, atol=1e-14)
        A.axpby_cols(al, A, be, B)                      # aliasing dst = A
        np.testing.assert_allclose(A.to_host(), a * al[None, :] + b * be[None, :], rtol=1e-14, atol=1e-14)
        B.remove_const_field()
        got = B.to_host()
        np.testing.assert_allclose(got, b - b.mean(axis=0)[None, :], rtol=1e-12, atol=1e-13)
        assert np.abs(got.sum(axis=0)).max() < 1e-9
        f = mx.MxMultiVector(m, 1, cx)
        frac = (np.arange(n) % 3 != 0).astype(np.float64) * 0.7
        f.from_host(frac.astype(np.complex128) if cx else frac)
        before = Dst.to_host()
        Dst.zero_unused(f)
        after = Dst.to_host()
        assert np.all(after[frac == 0] == 0) and np.array_equal(after[frac != 0], before[frac != 0])

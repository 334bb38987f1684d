"""CSG solids of the crab-cavity config C5 (reference example/crabcav.py:13-66; src/MxTorus.hpp, MxCone.hpp,
MxShapeUnion.hpp, MxShapeSubtract.hpp, MxShapeMirror.hpp, MxShapeRepeat.hpp, transforms MxShape.cpp:172-210)."""
import math
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_torus_and_cone_level_sets(orc):
    R, r = 0.3, 0.1
    tor = orc.Shape.torus(R, r, (0, 0, 1), (0.1, -0.2, 0.05))
    c = np.array([0.1, -0.2, 0.05])
    assert tor.func(c + (R, 0, 0)) == pytest.approx(r * r)               # centre of the tube
    assert tor.func(c + (R + r, 0, 0)) == pytest.approx(0.0, abs=1e-15)   # outer equator
    assert tor.func(c + (0, R, r)) == pytest.approx(0.0, abs=1e-15)       # top of the tube
    assert tor.func(c) < 0                                                # hole
    # quirk R14: the reference's in-plane gradient term has the opposite sign of d f / d rho
    g = tor.grad(c + (R + 0.5 * r, 0, 0))
    assert g[0] > 0 and g[1] == 0 and g[2] == 0
    h = 1e-6
    fd = (tor.func(c + (R + 0.5 * r + h, 0, 0)) - tor.func(c + (R + 0.5 * r - h, 0, 0))) / (2 * h)
    assert fd == pytest.approx(-g[0], rel=1e-6)
    gz = tor.grad(c + (R, 0, 0.5 * r))
    assert gz[2] == pytest.approx(-r)                                     # axial term is the true derivative
    th = 0.5
    cone = orc.Shape.cone(th, (0, 0, 2), (0, 0, 1.0))
    assert cone.func((0, 0, 3.0)) > 0 and cone.func((0, 0, -1.0)) > 0     # both nappes
    assert cone.func((2 * math.tan(th), 0, 3.0)) == pytest.approx(0.0, abs=1e-14)
    assert cone.func((1.0, 0, 1.0)) < 0
    gc = cone.grad((0.3, 0.0, 2.0))
    fdx = (cone.func((0.3 + h, 0, 2.0)) - cone.func((0.3 - h, 0, 2.0))) / (2 * h)
    fdz = (cone.func((0.3, 0, 2.0 + h)) - cone.func((0.3, 0, 2.0 - h))) / (2 * h)
    assert gc[0] == pytest.approx(fdx, rel=1e-6) and gc[2] == pytest.approx(fdz, rel=1e-6)


def test_torus_volume_from_cut_cells(orc):
    R, r, n = 0.3, 0.1, 40
    tor = orc.Shape.torus(R, r, (0, 0, 1), (0, 0, 0))
    h = 1.0 / n
    vol = 0.0
    for i in range(n):
        for j in range(n):
            for k in range(int(0.35 * n), int(0.65 * n)):
                vol += tor.fraction(2, 0, (h, h, h), (-0.5 + (i + 0.5) * h, -0.5 + (j + 0.5) * h, -0.5 + (k + 0.5) * h))
    vol *= h ** 3
    assert vol == pytest.approx(2 * math.pi ** 2 * R * r * r, rel=0.02)   # r = 4 cells


def test_boolean_mirror_repeat_semantics(orc):
    S = orc.Shape
    a, b = S.sphere(0.5, (0, 0, 0)), S.sphere(0.5, (0.6, 0, 0))
    u, d = S.union([a, b]), S.subtract(a, b)
    for p in [(0, 0, 0), (0.6, 0, 0), (0.3, 0, 0), (-0.6, 0, 0), (1.2, 0, 0), (0.05, 0.45, 0)]:
        fa, fb = a.func(p), b.func(p)
        assert u.func(p) == max(fa, fb)
        assert (d.func(p) > 0) == (fa > 0 and fb <= 0)
        if fa > 0 and fb > 0:
            assert d.func(p) == -fb                     # inside both: minus the removal shape's value
        elif fa <= 0 and fb <= 0:
            assert d.func(p) == fa
        else:
            assert d.func(p) == min(fa, -fb)
    assert np.array_equal(u.grad((0.5, 0.1, 0)), b.grad((0.5, 0.1, 0)))
    # mirror: positive side evaluates the shape, the other side its reflection
    hs = S.sphere(0.3, (0.1, 0.0, 0.2))
    m = S.mirror(hs, (0, 0, 3), (0, 0, 0))
    for p in [(0.1, 0.0, 0.2), (0.2, 0.1, 0.4), (0.0, 0.0, 0.05)]:
        q = (p[0], p[1], -p[2])
        assert m.func(p) == hs.func(p)
        assert m.func(q) == hs.func(p)
        assert np.array_equal(m.grad(q), hs.grad(p) * (1, 1, -1))
    # a reflected copy is its own inverse transform
    t = S.sphere(0.3, (0.1, 0.0, 0.2)).reflect((0, 0, 1), (0, 0, 0.1))
    assert t.func((0.1, 0.0, 0.0)) == pytest.approx(1.0)
    # repeat: period folding clamped to [-neg, pos] periods
    cell = S.sphere(0.2, (0, 0, 0))
    rep = S.repeat(cell, (0, 0, 0), (0, 0, 1), 0.5, 2, 1)
    assert rep.func((0, 0, 0.5)) == pytest.approx(1.0) and rep.func((0, 0, 1.0)) == pytest.approx(1.0)
    assert rep.func((0, 0, -0.5)) == pytest.approx(1.0)
    assert rep.func((0, 0, 1.5)) == pytest.approx(cell.func((0, 0, 0.5)))    # beyond the last positive copy
    assert rep.func((0, 0, -1.0)) == pytest.approx(cell.func((0, 0, -0.5)))
    assert rep.func((0.05, 0, 0.55)) == pytest.approx(cell.func((0.05, 0, 0.05)))


def test_crab_cavity_geometry_and_operators(orc):
    sh = orc.crabcav_shape()
    cell_len, cav_rad, iris_rad = 2 * 0.0192, 0.04719, 0.015
    assert sh.func((0, 0, 0)) > 0
    assert sh.func((cav_rad - 1e-4, 0, 0)) > 0 and sh.func((cav_rad + 1e-4, 0, 0)) < 0     # equator radius
    assert sh.func((iris_rad - 1e-4, 0, 0.5 * cell_len)) > 0 and sh.func((iris_rad + 1e-4, 0, 0.5 * cell_len)) < 0   # iris
    assert sh.func((0, 0, 2 * cell_len - 1e-4)) > 0 and sh.func((0, 0, 2 * cell_len + 1e-4)) < 0   # end caps
    rng = np.random.default_rng(0)
    for p in rng.uniform(-0.05, 0.05, (200, 3)):
        assert sh.func(p) == sh.func((p[0], p[1], -p[2]))                 # mirror symmetry
        q = (p[0], p[1], p[2] + cell_len)
        if abs(p[2]) < 0.4 * cell_len:
            assert sh.func(q) == pytest.approx(sh.func(p), abs=1e-15)     # periodic along z inside the caps
    sim = orc.crabcav(cell_res=8, pad=2)
    assert sim.n == (24, 24, 36)      # delta = cellLen / 8; nx = 2 (ceil(R / delta) + pad)
    gold = np.load(os.path.join(GOLD, "crabcav_counts.npz"))
    got = [len(sim.map(f)) for f in ("bfield", "efield", "psifield")]
    got += [sim.op(o).nnz for o in ("curlCurl", "gradDiv", "vecLapl", "scaLapl")]
    assert got == list(gold["counts"])
    np.testing.assert_allclose([sim.fracs(f).sum() for f in ("bfield", "efield", "psifield")], gold["frac_sums"], rtol=1e-13)
    fa, fv = sim.fracs("bfield"), sim.fracs("psifield")
    assert fa.min() >= 0 and fa.max() <= 1 + 1e-12 and 0 < (fa > 0).sum() < len(fa)
    cE, dB, gP = sim.op("curlE").scipy(), sim.op("divB").scipy(), sim.op("gradPsi").scipy()
    assert abs(dB @ cE).max() < 1e-6
    A = sim.op("curlCurl")
    assert abs(A.scipy() @ gP).max() < 1e-3 * abs(A.scipy()).max()
    assert np.diff(A.arrays()[0]).max() == 13
    # pi-mode-like cavity: lowest Maxwell eigenvalue is in the right range for a 47 mm equator radius
    # (TM010-like k ~ 2.405 / R ~ 51 1/m -> k^2 ~ 2600) -- coarse sanity bound only
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    L, M = sim.op("vecLapl").scipy(), sim.op("mRhs").scipy()
    d = M.diagonal()
    keep = np.where(d > 0)[0]
    w = sla.eigs(L[keep][:, keep].tocsc(), k=4, M=sp.diags(d[keep]).tocsc(), sigma=2000.0, tol=1e-8, return_eigenvectors=False)
    w = np.sort(w.real)
    assert 1500 < w[0] < 6000, w


def test_ellipsoid_rotate_scale(orc):
    """MxEllipsoid.hpp:31-50 and the rotate / scale placements of MxShape.cpp:89-168."""
    S = orc.Shape
    e = S.ellipsoid((0.1, 0.0, -0.2), (0.4, 0.2, 0.1))
    assert e.func((0.1, 0.0, -0.2)) == 1.0
    assert e.func((0.5, 0.0, -0.2)) == pytest.approx(0.0, abs=1e-15)
    assert e.func((0.1, 0.2, -0.2)) == pytest.approx(0.0, abs=1e-15) and e.func((0.1, 0.0, -0.1)) == pytest.approx(0.0, abs=1e-15)
    np.testing.assert_allclose(e.grad((0.3, 0.0, -0.2)), (-2 * 0.2 / 0.16, 0, 0), rtol=1e-13)
    # a scaled sphere has the same zero level set, hence the same cut-cell fractions
    s = S.sphere(0.1, (0, 0, 0)).scale((4.0, 2.0, 1.0))
    e0 = S.ellipsoid((0, 0, 0), (0.4, 0.2, 0.1))
    for p in [(0.38, 0.02, 0.01), (0.1, 0.17, 0.02), (0.05, 0.05, 0.09)]:
        for kind, axis, lens in ((0, 0, (0.05,)), (1, 2, (0.05, 0.05)), (2, 0, (0.05, 0.05, 0.05))):
            assert s.fraction(kind, axis, lens, p) == pytest.approx(e0.fraction(kind, axis, lens, p), abs=1e-10)
    # rotating a z-cylinder by a quarter turn about y (either sense) gives an x-cylinder
    for ang in (0.5 * math.pi, -0.5 * math.pi):
        c = S.cylinder(0.3, (0, 0, 1), (0, 0, 0)).rotate((0, 1, 0), ang)
        cx = S.cylinder(0.3, (1, 0, 0), (0, 0, 0))
        for p in [(5.0, 0.1, 0.2), (0.0, 0.3, 0.0), (-2.0, 0.2, 0.25), (0.0, 0.0, 0.31)]:
            assert c.func(p) == pytest.approx(cx.func(p), abs=1e-14)
    # sense of rotation as written in the reference (M = -[a]x): +angle about z takes the half-space normal x to -y
    h = S.halfspace((0, 0, 0), (1, 0, 0)).rotate((0, 0, 1), 0.5 * math.pi)
    assert h.func((0, -1, 0)) == pytest.approx(1.0) and h.func((1, 0, 0)) == pytest.approx(0.0, abs=1e-15)
    # rotation about an explicit pivot moves the centre; about the own centre (default pivot) it does not
    sp = S.sphere(0.2, (1.0, 0.0, 0.0)).rotate((0, 0, 1), math.pi, pivot=(0, 0, 0))
    assert sp.func((-1.0, 0.0, 0.0)) == pytest.approx(1.0)
    sp2 = S.sphere(0.2, (1.0, 0.0, 0.0)).rotate((0, 0, 1), 1.234)
    assert sp2.func((1.0, 0.0, 0.0)) == pytest.approx(1.0)
    # scaling about an origin moves the centre accordingly
    sc = S.sphere(0.1, (0.5, 0, 0)).scale((2.0, 2.0, 2.0), origin=(0.25, 0, 0))
    assert sc.func((0.75, 0, 0)) == pytest.approx(1.0) and sc.func((0.95, 0, 0)) == pytest.approx(0.0, abs=1e-14)


def test_iris_loaded_structure_with_bloch_phase(orc):
    """example/pillWTubes.py: pillbox + rounded iris tube (inverted tori), periodic in z with a 2 pi / 3 phase advance.
    Dey-Mittra PEC geometry AND complex Bloch factors in one operator; the structure is designed so that the TM010-like
    mode at that phase advance sits at 12 GHz."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    sim = orc.pill_w_tubes(cells_per_iris=2)
    assert sim.is_complex and sim.n == (28, 28, 10)
    A = sim.op("curlCurl")
    assert A.is_complex and np.diff(A.arrays()[0]).max() == 13
    dB, cE, gP = sim.op("divB").scipy(), sim.op("curlE").scipy(), sim.op("gradPsi").scipy()
    assert abs(dB @ cE).max() < 1e-9 * abs(cE).max()
    assert abs(A.scipy() @ gP).max() < 1e-9 * abs(A.scipy()).max() * abs(gP).max()
    L, M = sim.op("vecLapl").scipy(), sim.op("mRhs").scipy()
    d = M.diagonal().real
    keep = np.where(d > 0)[0]
    target = sim.info["k2_target"]                      # (2 pi 12 GHz / c)^2
    w = sla.eigs(L[keep][:, keep].tocsc(), k=4, M=sp.diags(d[keep]).tocsc(), sigma=target, tol=1e-9, return_eigenvectors=False)
    assert abs(w.imag).max() < 1e-6 * target
    nearest = w.real[np.argmin(abs(w.real - target))]
    assert abs(nearest - target) / target < 0.03, (nearest, target)
    # without the phase advance the operators are real
    assert not orc.pill_w_tubes(cells_per_iris=2, ph_adv=0.0).op("curlCurl").is_complex

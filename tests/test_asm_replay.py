"""CPU replay of the operator assembly against the oracle.

maxwell_b200/csrc/mxg_yee.h, mxg_shape.h, mxg_asm_impl.h and mxg_asm_api.inc hold the row generators, the CRS algebra,
the cut-cell fraction rules and the C entry points as host/device code. The product compiles them into CUDA kernels
(mxg_asm.cu); this test compiles the very same sources with plain loops (tests/cpp/asm_replay.cpp, symbols mxr_*) and
checks every map, fraction array and operator bit for bit against the oracle, so that the GPU run only has to confirm
that the device arithmetic rounds like the host's.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from maxwell_b200 import assembly as asm

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "maxwell_b200", "csrc")


@pytest.fixture(scope="module")
def api():
    out_dir = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libasm_replay.so")
    src = os.path.join(HERE, "cpp", "asm_replay.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in ("mxg_yee.h", "mxg_shape.h", "mxg_asm_impl.h", "mxg_asm_api.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall", "-Wno-sign-compare",
                               "-shared", "-I", CSRC, "-o", so, src])
    lib = C.CDLL(so)
    lib.mxr_last_error.restype = C.c_char_p
    return asm.AssemblyAPI(lib, "mxr_", lib.mxr_last_error)


# ---- the same geometry built twice: oracle shapes and product shapes -------------------------------------------------
def both(api, build):
    return build(orc.Shape), build(api)


pillbox_shape, crabcav_shape = asm.pillbox_shape, asm.crabcav_shape


def tilted_shape(S):
    """Exercises every placement operation and the remaining primitives."""
    e = S.ellipsoid((0.05, -0.02, 0.01), (0.33, 0.21, 0.27))
    e.rotate((1, 2, 3), 0.7)
    c = S.cylinder(0.15, (1, 1, 0), (0.1, 0.0, -0.05))
    c.rotate((0, 1, 0), -0.4, pivot=(0.2, 0.1, 0.0))
    s = S.sphere(0.2, (-0.15, 0.1, 0.05))
    s.scale((1.0, 0.5, 2.0), origin=(-0.1, 0.0, 0.0))
    h = S.halfspace((0.0, 0.0, 0.12), (0.1, -0.2, -1.0))
    t = S.torus(0.22, 0.06, (0, 1, 1), (0.0, 0.05, 0.0))
    t.reflect((1, 0, 0), (0.03, 0, 0))
    u = S.union([e, c, s, t])
    u.translate((0.01, 0.02, -0.03))
    k = S.cone(0.5, (0, 0, 1), (0.0, 0.0, -0.3))
    k.invert()
    return S.intersection([u, h, k])


SHAPES = {"pillbox": pillbox_shape, "crabcav": crabcav_shape, "tilted": tilted_shape}


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_shape_function_and_gradient_match_the_oracle(api, name):
    so, sp = both(api, SHAPES[name])
    rng = np.random.default_rng(7)
    scale = 0.09 if name == "crabcav" else 0.6
    pts = rng.uniform(-scale, scale, size=(400, 3))
    for p in pts:
        assert so.func(p) == sp.func(p)
        assert np.array_equal(np.asarray(so.grad(p)), sp.grad(p))


def make_pair(api, n, origin, size, shape=None, lower=None, upper=None, phase_shifts=None, literal=False, use_host_fracs=False):
    so = sp = None
    if shape is not None:
        so, sp = both(api, shape)
    o = orc.Sim(n, origin=origin, size=size, lower=lower, upper=upper, phase_shifts=phase_shifts, pec=so,
                literal_upper_periodic_e=literal)
    p = api.sim(None, n, origin=origin, size=size, lower=lower, upper=upper, phase_shifts=phase_shifts,
                literal_upper_periodic_e=literal)
    if shape is not None:
        if use_host_fracs:
            for f in asm.FIELDS:
                p.set_pec_fractions(f, o.full_fracs(f))
        else:
            p.set_pec_shape(sp)
    p.setup()
    return o, p


def crab_grid(cell_res=4):
    return asm.crabcav_grid(cell_res=cell_res)


CASES = {
    "vacuum": dict(n=6, origin=(0.0,) * 3, size=(1.0,) * 3),
    "vacuum-literal": dict(n=5, origin=(0.0,) * 3, size=(1.0,) * 3, literal=True),
    "vacuum-bloch": dict(n=(5, 6, 4), origin=(0.0,) * 3, size=(1.0, 1.2, 0.8), phase_shifts=(0.3, -0.7, 1.1)),
    "pillbox": dict(n=12, origin=(-0.5,) * 3, size=(1.0,) * 3, shape=pillbox_shape),
    "pillbox-hostfracs": dict(n=10, origin=(-0.5,) * 3, size=(1.0,) * 3, shape=pillbox_shape, use_host_fracs=True),
    "walls": dict(n=(6, 5, 7), origin=(0.0,) * 3, size=(0.5,) * 3, lower=(orc.PEC, orc.PMC, orc.PEC), upper=(orc.PEC,) * 3,
                  shape=lambda S: S.sphere(0.49, (0, 0, 0))),
    "walls-mixed": dict(n=6, origin=(0.0,) * 3, size=(1.0,) * 3, lower=(orc.PMC, orc.PERIODIC, orc.PEC),
                        upper=(orc.PEC, orc.PERIODIC, orc.PMC)),
    "tilted": dict(n=9, origin=(-0.5,) * 3, size=(1.0,) * 3, shape=tilted_shape),
    "crabcav": dict(zip(("n", "origin", "size"), crab_grid()), shape=crabcav_shape),
    "bloch-pec": dict(n=(8, 8, 6), origin=(-0.5, -0.5, 0.0), size=(1.0, 1.0, 0.7), phase_shifts=(0.0, 0.0, 2.0 * math.pi / 3.0),
                      shape=lambda S: S.union([S.intersection([S.cylinder(0.4, (0, 0, 1), (0, 0, 0)), S.slab(0.4, (0, 0, 1), (0, 0, 0.35))]),
                                               S.cylinder(0.15, (0, 0, 1), (0, 0, 0))])),
}
OPS = ("curlE", "curlB", "divB", "gradPsi", "dmA", "dmL", "dmVInv", "mRhs", "curlCurl", "gradDiv", "vecLapl", "scaLapl")


@pytest.mark.parametrize("case", sorted(CASES))
def test_maps_fractions_and_operators_match_the_oracle_bit_for_bit(api, case):
    o, p = make_pair(api, **CASES[case])
    for f in asm.FIELDS:
        assert np.array_equal(o.map(f), p.map(f)), f
        assert o.num_global(f) == p.num_global(f)
        ref = o.full_fracs(f)
        if ref is not None:
            assert np.array_equal(ref, p.fracs(f)), f
    for name in OPS:
        a, b = o.op(name), p.op(name)
        assert (a.nrows, a.ncols, a.nnz, a.is_complex) == (b.nrows, b.ncols, b.nnz, b.is_complex), name
        for x, y in zip(a.arrays(), b.arrays()):
            assert np.array_equal(x, y), name
    # real-valued build of a Bloch problem keeps the real parts (MxUtil.hpp:58-62)
    if CASES[case].get("phase_shifts"):
        for x, y in zip(o.op("curlCurl", is_complex=False).arrays(), p.op("curlCurl", is_complex=False).arrays()):
            assert np.array_equal(x, y)


def test_crs_algebra_entry_points(api):
    o, p = make_pair(api, **CASES["pillbox"])
    ce, cb, dl = p.op("curlE"), p.op("curlB"), p.op("dmL")
    cc = ce @ (dl @ cb)
    ref = o.op("curlCurl")
    for x, y in zip(ref.arrays(), cc.arrays()):
        assert np.array_equal(x, y)
    assert (cc.row_field, cc.col_field) == ("bfield", "bfield")
    gd = p.op("gradDiv")
    vl = cc.add(1.0, gd, -1.0, purge=True)
    for x, y in zip(o.op("vecLapl").arrays(), vl.arrays()):
        assert np.array_equal(x, y)
    unpurged = cc.add(1.0, gd, -1.0)
    assert unpurged.nnz >= vl.nnz
    for x, y in zip(vl.arrays(), unpurged.purge().arrays()):
        assert np.array_equal(x, y)
    ref_sum = o.op("curlCurl").add(2.0, o.op("gradDiv"), -0.5)
    for x, y in zip(ref_sum.arrays(), cc.add(2.0, gd, -0.5).arrays()):
        assert np.array_equal(x, y)
    # row slices come back rebased
    rp, col, val = cc.arrays(100, 180)
    frp, fcol, fval = cc.arrays()
    assert np.array_equal(rp, frp[100:181] - frp[100]) and np.array_equal(col, fcol[frp[100]:frp[180]])
    assert np.array_equal(val, fval[frp[100]:frp[180]])
    cc.scale(-2.0)
    assert np.array_equal(cc.arrays()[2], -2.0 * fval)
    with pytest.raises(asm.AssemblyError):
        ce @ ce      # shapes do not chain
    with pytest.raises(asm.AssemblyError):
        p.op("noSuchOperator")


def test_dielectric_chain_with_a_host_generated_inverse_permittivity(api):
    """MxYeeFitInvEps stays on the host; its matrix is uploaded and the chain around it runs here."""
    n = 8
    diel = orc.Shape.sphere(0.37, (0, 0, 0))
    o = orc.Sim(n, origin=(-0.5,) * 3, size=(1.0,) * 3, pec=orc.Shape.sphere(0.49, (0, 0, 0)), dielectrics=[(diel, orc.SAPPHIRE)])
    p = api.sim(None, n, origin=(-0.5,) * 3, size=(1.0,) * 3)
    p.set_pec_shape(api.sphere(0.49, (0, 0, 0))).setup()
    ie, iv = o.op("invEps"), o.op("invEpsVolAve")
    d_ie = p.upload("efield", "efield", *ie.arrays(), ncols=ie.ncols)
    d_iv = p.upload("psifield", "psifield", *iv.arrays(), ncols=iv.ncols)
    for name in ("curlCurl", "gradDiv", "vecLapl"):
        got = p.op(name, inv_eps=d_ie, inv_eps_vol_ave=d_iv)
        for x, y in zip(o.op(name).arrays(), got.arrays()):
            assert np.array_equal(x, y), name


@pytest.mark.parametrize("case", ["pillbox", "vacuum-bloch", "walls"])
def test_grid_transfers_match_the_oracle(api, case):
    """Prolongators (MxGridFieldInterpolator) between a grid and the one with half the cells, and the restrictions
    P^T / 8 the multigrid hierarchy of bench.py uses."""
    kw = dict(CASES[case])
    n = kw.pop("n")
    n = (n,) * 3 if np.isscalar(n) else n
    fine_n = tuple(2 * (v // 2) for v in n)
    coarse_n = tuple(v // 2 for v in fine_n)
    of, pf = make_pair(api, n=fine_n, **kw)
    oc, pc = make_pair(api, n=coarse_n, **kw)
    cx = bool(kw.get("phase_shifts"))
    for field in ("bfield", "psifield"):
        ref = orc.interpolator(oc, of, field=field, is_complex=cx)
        got = pf.interpolator_from(pc, field, is_complex=cx)
        assert (ref.nrows, ref.ncols, ref.nnz) == (got.nrows, got.ncols, got.nnz)
        for x, y in zip(ref.arrays(), got.arrays()):
            assert np.array_equal(x, y), field
        rt, gt = ref.transpose(scale=0.125), got.transpose(scale=0.125)
        assert (rt.nrows, rt.ncols, rt.nnz) == (gt.nrows, gt.ncols, gt.nnz)
        for x, y in zip(rt.arrays(), gt.arrays()):
            assert np.array_equal(x, y), field


DIELECTRIC_CASES = {
    # example/dsphmsph.py: dielectric sphere eps = 10 inside a PEC sphere
    "dsphmsph": dict(n=8, origin=(-0.5,) * 3, size=(1.0,) * 3, pec=lambda S: S.sphere(0.49, (0, 0, 0)),
                     diels=[(lambda S: S.sphere(0.37, (0, 0, 0)), np.eye(3) * 10.0)]),
    # example/run.py:15-28: the octant with PEC / PMC symmetry planes
    "dsphmsph-octant": dict(n=7, origin=(0.0,) * 3, size=(0.5,) * 3, lower=(orc.PEC, orc.PMC, orc.PMC), upper=(orc.PEC,) * 3,
                            pec=lambda S: S.sphere(0.49, (0, 0, 0)), diels=[(lambda S: S.sphere(0.37, (0, 0, 0)), np.eye(3) * 10.0)]),
    # example/phc-sapph-r0.37.py: anisotropic sapphire sphere, periodic with Bloch phases (complex)
    "phc-sapphire": dict(n=7, origin=(-0.5,) * 3, size=(1.0,) * 3, phase_shifts=(0.9, -0.4, 1.7),
                         diels=[(lambda S: S.sphere(0.37, (0, 0, 0)), orc.SAPPHIRE)]),
    # two objects, one with a complex (lossy) tensor, next to a PEC pillbox
    "two-objects": dict(n=8, origin=(-0.5,) * 3, size=(1.0,) * 3, pec=pillbox_shape,
                        diels=[(lambda S: S.sphere(0.2, (0.05, 0.0, -0.1)), orc.SAPPHIRE * (1.0 + 0.02j)),
                               (lambda S: S.ellipsoid((-0.1, 0.1, 0.15), (0.12, 0.2, 0.1)), np.diag([2.0, 3.0, 4.5]))]),
}


def dielectric_pair(new_sim, api, n, origin, size, diels, pec=None, lower=None, upper=None, phase_shifts=None):
    o = orc.Sim(n, origin=origin, size=size, lower=lower, upper=upper, phase_shifts=phase_shifts,
                pec=pec(orc.Shape) if pec else None, dielectrics=[(shape(orc.Shape), eps) for shape, eps in diels])
    p = new_sim(n, origin=origin, size=size, lower=lower, upper=upper, phase_shifts=phase_shifts)
    if pec:
        p.set_pec_shape(pec(api))
    for shape, eps in diels:
        p.add_dielectric(shape(api), eps)
    p.setup()
    return o, p


@pytest.mark.parametrize("case", sorted(DIELECTRIC_CASES))
def test_inverse_permittivity_and_dielectric_chains_match_the_oracle(api, case):
    """MxYeeFitInvEps: 9-point anisotropic rows from the edge / dual-face fractions and the interface normal, the
    cell-averaged scalar 1 / eps, and the chains that contain them."""
    kw = DIELECTRIC_CASES[case]
    o, p = dielectric_pair(lambda n, **k: api.sim(None, n, **k), api, **kw)
    cx = bool(kw.get("phase_shifts")) or any(np.iscomplexobj(e) for _, e in kw["diels"])
    for f in asm.FIELDS:
        assert np.array_equal(o.map(f), p.map(f)), f
    for name in ("invEps", "invEpsVolAve", "curlCurl", "gradDiv", "vecLapl"):
        a, b = o.op(name, is_complex=cx), p.op(name, is_complex=cx)
        assert (a.nrows, a.ncols, a.nnz) == (b.nrows, b.ncols, b.nnz), name
        for x, y in zip(a.arrays(), b.arrays()):
            assert np.array_equal(x, y), name
    if cx and not any(np.iscomplexobj(e) for _, e in kw["diels"]):
        for x, y in zip(o.op("invEps", is_complex=False).arrays(), p.op("invEps", is_complex=False).arrays()):
            assert np.array_equal(x, y)


def random_shape_builder(seed):
    """A seeded random CSG tree (primitives, placements, composites up to depth 4) as a function of the shape factory, so the
    oracle and the product build the very same solid."""
    def build(S):
        rng = np.random.default_rng(seed)

        def vec(scale=1.0):
            return tuple(float(v) for v in rng.uniform(-scale, scale, 3))

        def axis():
            v = rng.normal(size=3)
            return tuple(float(x) for x in v / np.linalg.norm(v))

        def primitive():
            k = int(rng.integers(0, 7))
            if k == 0:
                return S.cylinder(float(rng.uniform(0.1, 0.3)), axis(), vec(0.2))
            if k == 1:
                return S.sphere(float(rng.uniform(0.15, 0.35)), vec(0.2))
            if k == 2:
                return S.halfspace(vec(0.1), axis())
            if k == 3:
                return S.slab(float(rng.uniform(0.2, 0.5)), axis(), vec(0.1))
            if k == 4:
                return S.ellipsoid(vec(0.15), tuple(float(v) for v in rng.uniform(0.1, 0.35, 3)))
            if k == 5:
                return S.torus(float(rng.uniform(0.2, 0.3)), float(rng.uniform(0.05, 0.1)), axis(), vec(0.1))
            return S.cone(float(rng.uniform(0.3, 0.8)), axis(), vec(0.3))

        def place(s):
            k = int(rng.integers(0, 6))
            if k == 0:
                s.translate(vec(0.1))
            elif k == 1:
                s.rotate(axis(), float(rng.uniform(-1.5, 1.5)))
            elif k == 2:
                s.rotate(axis(), float(rng.uniform(-1.5, 1.5)), pivot=vec(0.2))
            elif k == 3:
                s.scale(tuple(float(v) for v in rng.uniform(0.7, 1.4, 3)), origin=vec(0.1))
            elif k == 4:
                s.reflect(axis(), vec(0.1))
            return s

        def tree(depth):
            if depth == 0 or rng.uniform() < 0.25:
                return place(primitive())
            k = int(rng.integers(0, 5))
            if k == 0:
                return place(S.intersection([tree(depth - 1), tree(depth - 1)]))
            if k == 1:
                return place(S.union([tree(depth - 1), tree(depth - 1), tree(depth - 1)]))
            if k == 2:
                return place(S.subtract(tree(depth - 1), [tree(depth - 1)]))
            if k == 3:
                return S.mirror(tree(depth - 1), axis(), vec(0.05))
            return S.repeat(tree(depth - 1), vec(0.05), axis(), float(rng.uniform(0.3, 0.6)), 1, 1)

        return tree(3)
    return build


@pytest.mark.parametrize("seed", range(8))
def test_random_csg_trees_fractions_and_operators(api, seed):
    """Eight random CSG solids: implicit function, gradient, edge / face / cell fractions, maps and the curl-curl chain agree
    with the oracle to the last bit."""
    build = random_shape_builder(1000 + seed)
    so, sp = both(api, build)
    rng = np.random.default_rng(seed)
    for p in rng.uniform(-0.5, 0.5, size=(200, 3)):
        assert so.func(p) == sp.func(p)
        assert np.array_equal(np.asarray(so.grad(p)), sp.grad(p))
    o, p = make_pair(api, n=(7, 6, 8), origin=(-0.5,) * 3, size=(1.0,) * 3, shape=build)
    for f in asm.FIELDS:
        assert np.array_equal(o.full_fracs(f), p.fracs(f)), f
        assert np.array_equal(o.map(f), p.map(f)), f
    if len(o.map("bfield")) > 0:
        for name in ("curlCurl", "vecLapl", "scaLapl"):
            for x, y in zip(o.op(name).arrays(), p.op(name).arrays()):
                assert np.array_equal(x, y), name


def test_too_deep_shapes_are_rejected(api):
    s = api.sphere(0.3, (0, 0, 0))
    for _ in range(8):
        s = api.intersection([s, api.halfspace((0, 0, 0.2), (0, 0, -1))])
    p = api.sim(None, 4, origin=(-0.5,) * 3, size=(1.0,) * 3)
    with pytest.raises(asm.AssemblyError):
        p.set_pec_shape(s)
    with pytest.raises(asm.AssemblyError):
        s.func((0, 0, 0))


def test_crs_algebra_on_general_matrices_against_scipy(api):
    """multiply / add / transpose / purge on random sparse matrices that are no stencils (rows of 0 .. 40 entries, products
    with up to a few hundred terms per row): pattern equal to scipy's, values to rounding."""
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    sim = api.sim(None, 2)

    def rand(m, n, density):
        a = sp.random(m, n, density=density, random_state=rng, format="csr", dtype=np.float64)
        a.sort_indices()
        return a

    def up(a):
        return sim.upload(None, None, a.indptr.astype(np.int64), a.indices.astype(np.int32), a.data, ncols=a.shape[1])

    def same(d, ref, exact_values=False):
        ref = ref.tocsr()
        ref.sort_indices()
        rowptr, col, val = d.arrays()
        assert np.array_equal(rowptr, ref.indptr) and np.array_equal(col, ref.indices)
        if exact_values:
            assert np.array_equal(val, ref.data)
        else:
            np.testing.assert_allclose(val, ref.data, rtol=1e-13, atol=1e-15)

    A, B, C = rand(300, 200, 0.05), rand(200, 250, 0.08), rand(300, 200, 0.03)
    dA, dB, dC = up(A), up(B), up(C)
    same(dA @ dB, A @ B)
    same(dA.add(2.0, dC, -0.5), 2.0 * A - 0.5 * C)
    same(dA.transpose(), A.T, exact_values=True)
    same((dA @ dB).transpose(scale=0.125), 0.125 * (A @ B).T)
    small = A.copy()
    small.data[::3] = 1e-13
    kept = small.copy()
    kept.data[np.abs(kept.data) <= 1e-12] = 0.0
    kept.eliminate_zeros()
    same(up(small).purge(), kept, exact_values=True)
    wide_a, wide_b = rand(80, 60, 0.08), rand(60, 400, 0.04)      # ~10 x ~25 terms per row: beyond the 160-entry row table
    same(up(wide_a) @ up(wide_b), wide_a @ wide_b)
    with pytest.raises(asm.AssemblyError):                        # the stated limit: no product row beyond 400 terms
        up(rand(50, 60, 0.6)) @ up(rand(60, 500, 0.5))
    # complex
    Z = rand(120, 90, 0.1).astype(np.complex128)
    Z.data = Z.data + 1j * rng.uniform(-1, 1, Z.nnz)
    Y = rand(90, 70, 0.1).astype(np.complex128)
    Y.data = Y.data * (0.3 - 0.8j)
    same(up(Z) @ up(Y), Z @ Y)
    same(up(Z).transpose(), Z.conj().T, exact_values=True)


def check_against_independent_fixtures(new_sim):
    """The product's assembly against fixtures that share no code with the oracle (tests/golden/make_golden.py): the scipy
    Kronecker-product curl-curl of the periodic vacuum box (DOF ids and pattern exact) and the analytic vacuum spectrum."""
    gold = os.path.join(HERE, "golden")
    g = np.load(os.path.join(gold, "vacuum_curlcurl_n6.npz"))
    p = new_sim(6, origin=(0.0,) * 3, size=(1.0,) * 3).setup()
    rowptr, col, val = p.op("curlCurl").arrays()
    assert np.array_equal(p.map("bfield"), g["gids"])
    assert np.array_equal(rowptr, g["indptr"]) and np.array_equal(col, g["indices"])
    np.testing.assert_allclose(val, g["data"], rtol=1e-14, atol=1e-12)
    import scipy.sparse as sp
    n = 8
    q = new_sim(n, origin=(0.0,) * 3, size=(1.0,) * 3).setup()
    rp, c, v = q.op("vecLapl").arrays()
    w = np.linalg.eigvalsh(sp.csr_matrix((v, c, rp)).toarray())
    spec = np.load(os.path.join(gold, "vacuum_spectrum.npz"))["n8"]
    np.testing.assert_allclose(w[:40], np.repeat(spec, 3)[:40], rtol=1e-10, atol=1e-9)
    # structural invariants the reference itself probes (MxMagWaveOp.cpp:644-653): div curl = 0, curlCurl grad = 0
    o = new_sim(10, origin=(-0.5,) * 3, size=(1.0,) * 3)
    o.set_pec_shape(pillbox_shape(o.api)).setup()
    div_curl = o.op("divB") @ o.op("curlE")
    assert np.abs(div_curl.arrays()[2]).max() < 1e-9
    cc_grad = o.op("curlCurl") @ o.op("gradPsi")
    assert np.abs(cc_grad.arrays()[2]).max() < 1e-7
    cc = o.op("curlCurl")
    assert np.diff(cc.arrays()[0]).max() == 13            # bulk rows of curl-curl have 13 entries


def test_assembly_against_fixtures_independent_of_the_oracle(api):
    check_against_independent_fixtures(lambda n, **k: api.sim(None, n, **k))

"""GPU parity on the dielectric configs: C3 (dielectric sphere in metal sphere, 9-point eps^-1 rows, up to 33 nnz
per curl-curl row) and C4 (sapphire photonic crystal with Bloch phases: complex128, K-form order)."""
import numpy as np
import pytest

from conftest import gpu_matrix, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", [0, 1])
def test_dielectric_sphere_curlcurl_bit_exact(mx, ctx, orc, layout):
    sim = orc.dsphmsph(20)
    A, op, rmap, _ = gpu_matrix(mx, ctx, sim, "curlCurl", layout=layout)
    assert np.diff(op.arrays()[0]).max() > 13
    x = mx.MxMultiVector(rmap, 3)
    y = mx.MxMultiVector(rmap, 3)
    x.random(31)
    A.apply(x, y)
    assert np.array_equal(op.apply(x.to_host()), y.to_host())


def test_dielectric_operators_on_the_path(mx, ctx, orc):
    sim = orc.dsphmsph(16)
    for name in ("vecLapl", "gradDiv", "scaLapl"):
        A, op, rmap, cmap = gpu_matrix(mx, ctx, sim, name)
        x = mx.MxMultiVector(cmap, 2)
        y = mx.MxMultiVector(rmap, 2)
        x.random(7)
        A.apply(x, y)
        assert np.array_equal(op.apply(x.to_host()), y.to_host()), name


@pytest.mark.parametrize("layout", [0, 1])
def test_sapphire_crystal_complex_kform_bitwise(mx, ctx, orc, layout):
    sim = orc.phc_sapphire(12, phase_shifts=(0.9, -0.4, 0.25))
    A, op, rmap, _ = gpu_matrix(mx, ctx, sim, "curlCurl", layout=layout)
    assert op.is_complex
    x = mx.MxMultiVector(rmap, 2, True)
    y = mx.MxMultiVector(rmap, 2, True)
    x.random(5)
    A.apply(x, y)
    got, xh = y.to_host(), x.to_host()
    assert rel_err(got, op.apply(xh)) < 1e-14
    K = op.kform()
    for j in range(2):
        yk = K.apply(np.ascontiguousarray(xh[:, j]).view(np.float64)).view(np.complex128)
        assert np.array_equal(yk, got[:, j])


@pytest.mark.parametrize("pol", ["TM", "TE"])
def test_c3_octant_with_symmetry_planes_bit_exact(mx, ctx, orc, pol):
    """Config C3 as the reference runs it: one octant, PEC/PMC symmetry planes per multipole family
    (example/run.py:169-179, example/dsphmsph.py:436-493). The mirror-boundary rows make the operators non-symmetric;
    parity is the bit-exact apply."""
    sim = orc.dsphmsph_octant(12, orc.oct_lower_bcs(pol, 1, 0))
    for name in ("curlCurl", "vecLapl"):
        A, op, rmap, _ = gpu_matrix(mx, ctx, sim, name)
        x = mx.MxMultiVector(rmap, 3)
        y = mx.MxMultiVector(rmap, 3)
        x.random(11)
        A.apply(x, y)
        assert np.array_equal(op.apply(x.to_host()), y.to_host()), (pol, name)

"""GPU parity on config C5 (crab cavity CSG, reference example/crabcav.py) at a test-sized grid: the assembled
operators of the eigensolve path applied on the GPU are bit-identical to the Epetra-order CPU apply, in both layouts.
The example puts the end caps exactly on a grid plane, so with the default Dey-Mittra fraction 0 the inverse volume
fractions reach ~1e19 (as they do in the reference); the parity bar is bit-exactness regardless of scaling."""
import numpy as np
import pytest

from conftest import gpu_matrix

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", [0, 1])
def test_crab_cavity_operators_bit_exact(mx, ctx, orc, layout):
    sim = orc.crabcav(cell_res=8, pad=2)
    for name in ("curlCurl", "vecLapl", "divB", "gradPsi", "scaLapl"):
        A, op, rmap, cmap = gpu_matrix(mx, ctx, sim, name, layout=layout)
        x = mx.MxMultiVector(cmap, 4)
        y = mx.MxMultiVector(rmap, 4)
        x.random(101)
        A.apply(x, y)
        assert np.array_equal(op.apply(x.to_host()), y.to_host()), name
    st = A.stats()
    assert st["nnz"] == op.nnz


def test_iris_loaded_structure_complex_parity(mx, ctx, orc):
    """example/pillWTubes.py at a test-sized grid: Dey-Mittra cut cells and Bloch phase factors in the same complex
    operator. Parity: within 1e-14 of the complex CSR apply and bit-identical to the reference's real 2N K-form order."""
    from conftest import rel_err
    sim = orc.pill_w_tubes(cells_per_iris=2)
    for name in ("curlCurl", "vecLapl"):
        A, op, rmap, _ = gpu_matrix(mx, ctx, sim, name)
        assert op.is_complex
        x = mx.MxMultiVector(rmap, 2, True)
        y = mx.MxMultiVector(rmap, 2, True)
        x.random(9)
        A.apply(x, y)
        got, xh = y.to_host(), x.to_host()
        assert rel_err(got, op.apply(xh)) < 1e-14, name
        K = op.kform()
        for j in range(2):
            yk = K.apply(np.ascontiguousarray(xh[:, j]).view(np.float64)).view(np.complex128)
            assert np.array_equal(yk, got[:, j]), name

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

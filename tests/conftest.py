import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def mx():
    import __graft_entry__ as ge
    ge.build_lib()
    import maxwell_b200
    maxwell_b200.load_library()
    return maxwell_b200


@pytest.fixture(scope="session")
def ctx(mx):
    """GPU context; constructing it fails loudly when no device / library is present."""
    return mx.Context(0)


def gpu_matrix(mx, ctx, sim, name, layout=None, is_complex=None):
    """Upload an oracle-generated operator through the C ABI; returns (A, op, row_map, col_map)."""
    op = sim.op(name, is_complex)
    rowptr, col, val = op.arrays()
    rg, cg = op.maps()
    fields = {"curlCurl": ("bfield", "bfield"), "vecLapl": ("bfield", "bfield"), "gradDiv": ("bfield", "bfield"),
              "mRhs": ("bfield", "bfield"), "dmA": ("bfield", "bfield"), "curlE": ("bfield", "efield"),
              "curlB": ("efield", "bfield"), "divB": ("psifield", "bfield"), "gradPsi": ("bfield", "psifield"),
              "scaLapl": ("psifield", "psifield"), "invEps": ("efield", "dfield")}[name]
    rmap = mx.MxMap(ctx, sim.num_global(fields[0]), rg)
    cmap = rmap if fields[0] == fields[1] else mx.MxMap(ctx, sim.num_global(fields[1]), cg)
    A = mx.MxCrsMatrix.from_csr(rmap, cmap, rowptr, cg[col], val, layout=layout)
    return A, op, rmap, cmap


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)

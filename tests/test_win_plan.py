"""Host-side tile planner of the windowed dictionary SpMM (maxwell_b200/csrc/mxg_spmm_win.cuh). CPU only: the planner runs
on the pattern tables of real operators from the oracle and a C++ replay of the kernel's shared-memory addressing checks
that every (row, entry) finds its column inside the staged windows. The GPU kernel itself is covered bit-for-bit by
tests/test_gpu_spmv.py (every parity test runs through it)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_ilv_model import _pattern_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
INC = os.path.join(ROOT, "maxwell_b200", "csrc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("win") / "libwin.so")
    subprocess.check_call([CXX, "-std=c++17", "-O2", "-shared", "-fPIC", "-Wall", "-Werror", "-I", INC,
                           os.path.join(ROOT, "tests", "cpp", "win_plan_capi.cpp"), "-o", so])
    return C.CDLL(so)


def _plan(lib, op, R, align, budget):
    rp, po, d = _pattern_table(op)
    out = (C.c_int64 * 6)()
    vp = C.c_void_p
    lib.win_plan_eval.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, vp]
    rc = lib.win_plan_eval(rp.ctypes.data, po.ctypes.data, len(po) - 1, d.ctypes.data, op.nrows, op.ncols, R, align, budget, out)
    assert rc == 0
    keys = ["tiles", "valid", "max_total", "bad", "rows_windowed", "copied"]
    return dict(zip(keys, list(out))), int((rp >= 0).sum())


@pytest.mark.parametrize("name", ["curlCurl", "vecLapl", "scaLapl"])
def test_windows_cover_every_gather_pillbox(lib, orc, name):
    op = orc.pillbox(32).op(name)
    for R, align in ((1536, 2), (768, 1), (384, 2)):
        res, dict_rows = _plan(lib, op, R, align, 12800)
        assert res["bad"] == 0, res
        assert res["valid"] == res["tiles"], res                     # a 32^3 grid fits the budget everywhere
        assert res["rows_windowed"] == dict_rows
        # three windows of about R rows plus the +-y lines: far below "every row loads its whole stencil"
        assert res["max_total"] <= 3 * R + 4 * 3 * 34 + 64


def test_windows_on_periodic_vacuum_and_dielectric(lib, orc):
    for sim, name in ((orc.vacuum(16), "curlCurl"), (orc.dsphmsph(16), "curlCurl"), (orc.vacuum(12, phase_shifts=(0.3, 0.2, 0.1)), "vecLapl")):
        res, dict_rows = _plan(lib, sim.op(name), 1536, 2, 12800)
        assert res["bad"] == 0, res
        assert res["rows_windowed"] <= dict_rows


def test_budget_overflow_falls_back(lib, orc):
    op = orc.pillbox(24).op("curlCurl")
    res, _ = _plan(lib, op, 1536, 2, 2000)    # smaller than one tile's own rows: nothing can be windowed
    assert res["valid"] == 0 and res["bad"] == 0

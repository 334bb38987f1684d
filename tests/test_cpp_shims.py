"""The C++ shim classes (include/mx/*.hpp: MxComm, MxMap, MxMultiVector, MxAnasaziMV, MxCrsMatrix, MxSolver) compile
against the C ABI with a plain host compiler and behave like the reference's test/AnasaziInterface.cpp expects."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, mx):
    exe = str(tmp_path / "shim_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(mx.library_path())
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_check.cpp"), "-o", exe, "-L", libdir, "-lmxgpu",
                           "-Wl,-rpath," + libdir])
    return exe


def test_shims_compile_and_fail_loudly_without_gpu(tmp_path, mx):
    exe = _build(tmp_path, mx)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if res.returncode == 0:
        assert "PASSED" in res.stdout          # a GPU is present: the checks themselves ran
    else:
        assert res.returncode == 3 and "no CPU fallback" in res.stdout, res.stdout + res.stderr


@pytest.mark.gpu
def test_shims_on_gpu(tmp_path, mx, ctx):
    exe = _build(tmp_path, mx)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PASSED" in res.stdout, res.stdout + res.stderr


def test_dense_helpers_on_cpu(tmp_path, mx):
    """mx::dense (Cholesky, Jacobi symEig, generalized / rank-revealing generalized eigenproblem): host logic of the
    LOBPCG driver, checked without a GPU."""
    exe = str(tmp_path / "dense_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(mx.library_path())
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "dense_check.cpp"), "-o", exe, "-L", libdir, "-lmxgpu",
                           "-Wl,-rpath," + libdir])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "PASSED" in res.stdout, res.stdout + res.stderr


def test_public_headers_are_plain_c(tmp_path, mx):
    exe = str(tmp_path / "abi_is_c")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    libdir = os.path.dirname(mx.library_path())
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "abi_is_c.c"), "-o", exe, "-L", libdir, "-lmxgpu", "-lmxsolver",
                           "-Wl,-rpath," + libdir])
    assert subprocess.run([exe], timeout=60).returncode == 0


def test_solver_driver_on_host_multivector(tmp_path, mx):
    """MxSolverT (the LOBPCG driver behind MxSolver) instantiated on a plain host multivector (tests/cpp/host_mv.hpp):
    eigenvalues of Dirichlet Laplacians against the analytic spectrum, degenerate eigenvalues, a generalized problem
    with a preconditioner, rejection of rank-deficient start blocks. Same code as the GPU instantiation."""
    exe = str(tmp_path / "solver_host_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(mx.library_path())
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp"),
                           os.path.join(ROOT, "tests", "cpp", "solver_host_check.cpp"), "-o", exe, "-L", libdir, "-lmxgpu",
                           "-Wl,-rpath," + libdir])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PASSED" in res.stdout, res.stdout + res.stderr


def test_assembly_shims(tmp_path, mx):
    """include/mx/MxAssembly.hpp (MxShape / MxEMSim / MxDeviceCrs over include/mxasm.h): shapes are host objects and are
    checked everywhere; without a device the simulation must fail loudly, with one the program assembles and applies a
    small curl-curl operator two ways."""
    exe = str(tmp_path / "asm_shim_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(mx.library_path())
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "asm_shim_check.cpp"), "-o", exe, "-L", libdir, "-lmxgpu",
                           "-Wl,-rpath," + libdir])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "PASSED" in res.stdout, res.stdout + res.stderr

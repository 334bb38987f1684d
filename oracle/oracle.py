"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package maxwell_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PERIODIC, ZERO, CONSTANT, PEC, PMC = 0, 1, 2, 3, 4


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, i64, dbl, cp = C.c_void_p, C.c_int64, C.c_double, C.c_char_p
        d3 = C.POINTER(C.c_double)
        i3 = C.POINTER(C.c_int)
        sig = {
            "mxo_last_error": (cp, []),
            "mxo_shape_cylinder": (vp, [dbl, d3, d3]),
            "mxo_shape_sphere": (vp, [dbl, d3]),
            "mxo_shape_slab": (vp, [dbl, d3, d3]),
            "mxo_shape_halfspace": (vp, [d3, d3]),
            "mxo_shape_intersection": (vp, [C.POINTER(vp), C.c_int]),
            "mxo_shape_torus": (vp, [dbl, dbl, d3, d3]),
            "mxo_shape_cone": (vp, [dbl, d3, d3]),
            "mxo_shape_union": (vp, [C.POINTER(vp), C.c_int]),
            "mxo_shape_subtract": (vp, [vp, C.POINTER(vp), C.c_int]),
            "mxo_shape_mirror": (vp, [vp, d3, d3]),
            "mxo_shape_repeat": (vp, [vp, d3, d3, dbl, C.c_int, C.c_int]),
            "mxo_shape_ellipsoid": (vp, [d3, d3]),
            "mxo_shape_rotate": (None, [vp, d3, dbl, vp]),
            "mxo_shape_scale": (None, [vp, d3, d3]),
            "mxo_shape_translate": (None, [vp, d3]),
            "mxo_shape_reflect": (None, [vp, d3, d3]),
            "mxo_shape_grad": (None, [vp, d3, d3]),
            "mxo_shape_invert": (None, [vp]),
            "mxo_shape_func": (dbl, [vp, d3]),
            "mxo_shape_destroy": (None, [vp]),
            "mxo_fraction": (dbl, [vp, C.c_int, C.c_int, d3, d3]),
            "mxo_sim_create": (vp, [i3, d3, d3]),
            "mxo_sim_destroy": (None, [vp]),
            "mxo_sim_set_bcs": (None, [vp, i3, i3]),
            "mxo_sim_set_phase_shifts": (None, [vp, d3]),
            "mxo_sim_set_pec": (None, [vp, vp]),
            "mxo_sim_add_dielectric": (None, [vp, vp, vp, cp]),
            "mxo_sim_set_literal_upper_periodic_e": (None, [vp, C.c_int]),
            "mxo_sim_setup": (C.c_int, [vp]),
            "mxo_sim_map_size": (i64, [vp, cp]),
            "mxo_sim_map_copy": (C.c_int, [vp, cp, vp]),
            "mxo_sim_map_num_global": (i64, [vp, cp]),
            "mxo_sim_rep_copy": (i64, [vp, cp, cp, vp]),
            "mxo_sim_map_fracs": (C.c_int, [vp, cp, vp]),
            "mxo_build_op": (vp, [vp, cp, C.c_int]),
            "mxo_mat_kform": (vp, [vp]),
            "mxo_mat_multiply": (vp, [vp, vp]),
            "mxo_build_interp": (vp, [vp, vp, cp, C.c_int]),
            "mxo_mat_transpose": (vp, [vp, C.c_int]),
            "mxo_mat_add": (vp, [vp, d3, vp, d3]),
            "mxo_mat_scale": (None, [vp, d3]),
            "mxo_mat_destroy": (None, [vp]),
            "mxo_mat_is_complex": (C.c_int, [vp]),
            "mxo_mat_shape": (None, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
            "mxo_mat_copy": (None, [vp, vp, vp, vp]),
            "mxo_mat_maps": (None, [vp, vp, vp]),
            "mxo_mat_apply": (C.c_int, [vp, vp, i64, vp, i64, C.c_int, C.c_int]),
            "mxo_num_threads": (C.c_int, []),
            "mxo_csr_apply": (C.c_int, [i64, vp, vp, vp, C.c_int, vp, i64, vp, i64, C.c_int, C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().mxo_last_error().decode())


class Shape:
    def __init__(self, handle, keep=()):
        self.h = handle
        self._keep = keep

    @staticmethod
    def cylinder(r, axis, loc):
        return Shape(lib().mxo_shape_cylinder(r, _d3(axis), _d3(loc)))

    @staticmethod
    def sphere(r, loc):
        return Shape(lib().mxo_shape_sphere(r, _d3(loc)))

    @staticmethod
    def slab(thickness, normal, loc):
        return Shape(lib().mxo_shape_slab(thickness, _d3(normal), _d3(loc)))

    @staticmethod
    def halfspace(point, normal):
        return Shape(lib().mxo_shape_halfspace(_d3(point), _d3(normal)))

    @staticmethod
    def intersection(shapes):
        arr = (C.c_void_p * len(shapes))(*[s.h for s in shapes])
        return Shape(lib().mxo_shape_intersection(arr, len(shapes)), keep=tuple(shapes))

    @staticmethod
    def torus(major_radius, minor_radius, axis, loc):
        return Shape(lib().mxo_shape_torus(major_radius, minor_radius, _d3(axis), _d3(loc)))

    @staticmethod
    def cone(angle, axis, vertex):
        return Shape(lib().mxo_shape_cone(angle, _d3(axis), _d3(vertex)))

    @staticmethod
    def union(shapes):
        arr = (C.c_void_p * len(shapes))(*[s.h for s in shapes])
        return Shape(lib().mxo_shape_union(arr, len(shapes)), keep=tuple(shapes))

    @staticmethod
    def subtract(base, removed):
        removed = list(removed) if isinstance(removed, (list, tuple)) else [removed]
        arr = (C.c_void_p * len(removed))(*[s.h for s in removed])
        return Shape(lib().mxo_shape_subtract(base.h, arr, len(removed)), keep=(base,) + tuple(removed))

    @staticmethod
    def mirror(shape, normal, point):
        return Shape(lib().mxo_shape_mirror(shape.h, _d3(normal), _d3(point)), keep=(shape,))

    @staticmethod
    def repeat(shape, origin, direction, step, num_pos, num_neg):
        return Shape(lib().mxo_shape_repeat(shape.h, _d3(origin), _d3(direction), float(step), int(num_pos), int(num_neg)),
                     keep=(shape,))

    @staticmethod
    def ellipsoid(loc, axes):
        return Shape(lib().mxo_shape_ellipsoid(_d3(loc), _d3(axes)))

    def rotate(self, axis, angle, pivot=None):
        """MxShape::rotate (MxShape.cpp:89-125): about `pivot`, or about the shape's current translation point."""
        pv = _d3(pivot) if pivot is not None else None
        lib().mxo_shape_rotate(self.h, _d3(axis), float(angle), C.cast(pv, C.c_void_p) if pv is not None else None)
        return self

    def scale(self, magnitudes, origin=(0.0, 0.0, 0.0)):
        lib().mxo_shape_scale(self.h, _d3(magnitudes), _d3(origin))
        return self

    def translate(self, v):
        lib().mxo_shape_translate(self.h, _d3(v))
        return self

    def reflect(self, normal, point):
        lib().mxo_shape_reflect(self.h, _d3(normal), _d3(point))
        return self

    def invert(self):
        lib().mxo_shape_invert(self.h)
        return self

    def func(self, p):
        return lib().mxo_shape_func(self.h, _d3(p))

    def grad(self, p):
        g = (C.c_double * 3)()
        lib().mxo_shape_grad(self.h, _d3(p), g)
        return np.array(list(g))

    def fraction(self, kind, axis, lens, p):
        return lib().mxo_fraction(self.h, kind, axis, _d3(list(lens) + [0.0] * (3 - len(lens))), _d3(p))

    def __del__(self):
        try:
            lib().mxo_shape_destroy(self.h)
        except Exception:
            pass


class Matrix:
    """CSR with local column indices; rows/cols follow the row/col map order."""

    def __init__(self, handle):
        self.h = handle
        L = lib()
        nr, nc, nz = C.c_int64(), C.c_int64(), C.c_int64()
        L.mxo_mat_shape(handle, C.byref(nr), C.byref(nc), C.byref(nz))
        self.nrows, self.ncols, self.nnz = nr.value, nc.value, nz.value
        self.is_complex = bool(L.mxo_mat_is_complex(handle))
        self._arrays = None

    def arrays(self):
        if self._arrays is None:
            rowptr = np.empty(self.nrows + 1, dtype=np.int64)
            col = np.empty(self.nnz, dtype=np.int32)
            val = np.empty(self.nnz, dtype=np.complex128 if self.is_complex else np.float64)
            lib().mxo_mat_copy(self.h, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data)
            self._arrays = (rowptr, col, val)
        return self._arrays

    def maps(self):
        rg = np.empty(self.nrows, dtype=np.int64)
        cg = np.empty(self.ncols, dtype=np.int64)
        lib().mxo_mat_maps(self.h, rg.ctypes.data, cg.ctypes.data)
        return rg, cg

    def scipy(self):
        import scipy.sparse as sp
        rowptr, col, val = self.arrays()
        return sp.csr_matrix((val, col, rowptr), shape=(self.nrows, self.ncols))

    def apply(self, X, nthreads=0):
        """Y = A X in the reference (Epetra) summation order. X: (ncols,) or (ncols, b) Fortran-ordered."""
        dt = np.complex128 if self.is_complex else np.float64
        X2 = np.asfortranarray(np.asarray(X, dtype=dt).reshape(self.ncols, -1))
        b = X2.shape[1]
        Y = np.zeros((self.nrows, b), dtype=dt, order="F")
        _check(lib().mxo_mat_apply(self.h, X2.ctypes.data, self.ncols, Y.ctypes.data, self.nrows, b, nthreads))
        return Y if np.ndim(X) == 2 else Y[:, 0]

    def kform(self):
        h = lib().mxo_mat_kform(self.h)
        if not h:
            raise RuntimeError(lib().mxo_last_error().decode())
        return Matrix(h)

    def transpose(self, normalize_rows=False, scale=None):
        """(conjugate) transpose, optionally row-normalised or multiplied by `scale`."""
        m = Matrix(lib().mxo_mat_transpose(self.h, int(normalize_rows)))
        if scale is not None:
            s = complex(scale)
            lib().mxo_mat_scale(m.h, (C.c_double * 2)(s.real, s.imag))
        return m

    def add(self, sa, other, sb):
        """sa*self + sb*other on the union pattern."""
        sa, sb = complex(sa), complex(sb)
        h = lib().mxo_mat_add(self.h, (C.c_double * 2)(sa.real, sa.imag), other.h, (C.c_double * 2)(sb.real, sb.imag))
        if not h:
            raise RuntimeError(lib().mxo_last_error().decode())
        return Matrix(h)

    def __matmul__(self, other):
        h = lib().mxo_mat_multiply(self.h, other.h)
        if not h:
            raise RuntimeError(lib().mxo_last_error().decode())
        return Matrix(h)

    def __del__(self):
        try:
            lib().mxo_mat_destroy(self.h)
        except Exception:
            pass


class Sim:
    """Mirror of MxEMSim (MxEMSim.cpp:54-225) restricted to 3-D, PEC shapes and BCs."""

    def __init__(self, n, origin=(0.0, 0.0, 0.0), size=(1.0, 1.0, 1.0), lower=None, upper=None,
                 phase_shifts=None, pec=None, literal_upper_periodic_e=False, dielectrics=()):
        if np.isscalar(n):
            n = (n, n, n)
        L = lib()
        self.n = tuple(int(x) for x in n)
        self.h = L.mxo_sim_create(_i3(n), _d3(origin), _d3(size))
        if lower is not None or upper is not None:
            L.mxo_sim_set_bcs(self.h, _i3(lower or (0, 0, 0)), _i3(upper or (0, 0, 0)))
        if phase_shifts is not None:
            L.mxo_sim_set_phase_shifts(self.h, _d3(phase_shifts))
        self.is_complex = phase_shifts is not None and any(p != 0 for p in phase_shifts)
        self._pec = pec
        if pec is not None:
            L.mxo_sim_set_pec(self.h, pec.h)
        L.mxo_sim_set_literal_upper_periodic_e(self.h, int(literal_upper_periodic_e))
        self._diels = list(dielectrics)
        for i, (shape, eps) in enumerate(self._diels):
            e = np.ascontiguousarray(np.asarray(eps, dtype=np.complex128).reshape(3, 3))
            L.mxo_sim_add_dielectric(self.h, shape.h, e.ctypes.data, ("diel%d" % i).encode())
        _check(L.mxo_sim_setup(self.h))

    def map(self, field):
        n = lib().mxo_sim_map_size(self.h, field.encode())
        if n < 0:
            raise RuntimeError(lib().mxo_last_error().decode())
        out = np.empty(n, dtype=np.int64)
        _check(lib().mxo_sim_map_copy(self.h, field.encode(), out.ctypes.data))
        return out

    def num_global(self, field):
        return lib().mxo_sim_map_num_global(self.h, field.encode())

    def fracs(self, field):
        n = lib().mxo_sim_map_size(self.h, field.encode())
        out = np.empty(n, dtype=np.float64)
        _check(lib().mxo_sim_map_fracs(self.h, field.encode(), out.ctypes.data))
        return out

    def full_fracs(self, field, rep="pec"):
        """Fractions of a shape representation on the guarded block ((N+3)^3 cells x components); None if absent."""
        n = lib().mxo_sim_rep_copy(self.h, field.encode(), rep.encode(), None)
        if n < 0:
            raise RuntimeError(lib().mxo_last_error().decode())
        if n == 0:
            return None
        out = np.empty(n, dtype=np.float64)
        lib().mxo_sim_rep_copy(self.h, field.encode(), rep.encode(), out.ctypes.data)
        return out

    def op(self, name, is_complex=None):
        cx = self.is_complex if is_complex is None else is_complex
        h = lib().mxo_build_op(self.h, name.encode(), int(cx))
        if not h:
            raise RuntimeError(lib().mxo_last_error().decode())
        return Matrix(h)

    def __del__(self):
        try:
            lib().mxo_sim_destroy(self.h)
        except Exception:
            pass


def interpolator(sim_from, sim_to, field="bfield", is_complex=False):
    """Trilinear interpolation of `field` from sim_from's grid to sim_to's component positions
    (MxGridFieldInterpolator.cpp:28-122): the refiners/coarseners of the multigrid spec."""
    h = lib().mxo_build_interp(sim_from.h, sim_to.h, field.encode(), int(is_complex))
    if not h:
        raise RuntimeError(lib().mxo_last_error().decode())
    return Matrix(h)


def csr_apply(rowptr, col, val, X, nthreads=0):
    """Epetra-order CSR apply on raw arrays (CPU baseline on arbitrary row blocks)."""
    cx = np.iscomplexobj(val)
    dt = np.complex128 if cx else np.float64
    nrows = len(rowptr) - 1
    X2 = np.asfortranarray(np.asarray(X, dtype=dt).reshape(len(X), -1))
    b = X2.shape[1]
    Y = np.zeros((nrows, b), dtype=dt, order="F")
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=dt)
    _check(lib().mxo_csr_apply(nrows, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, int(cx),
                               X2.ctypes.data, X2.shape[0], Y.ctypes.data, nrows, b, nthreads))
    return Y if np.ndim(X) == 2 else Y[:, 0]


def pillbox(n, radius=0.4, length=0.8, origin=-0.5, size=1.0):
    """example/pillbox.py:14-17 -- Cylinder(R, axis z) ∩ Slab(thickness, normal z) in a box."""
    cyl = Shape.cylinder(radius, (0, 0, 1), (0, 0, 0))
    caps = Shape.slab(length, (0, 0, 1), (0, 0, 0))
    cav = Shape.intersection([cyl, caps])
    return Sim(n, origin=(origin,) * 3, size=(size,) * 3, pec=cav)


def crabcav_shape(num_cells=4, cell_len=2.0 * 0.0192, cav_rad=0.04719, iris_rad=0.015, cav_rho=0.0136, iris_rho=0.00331):
    """example/crabcav.py:13-66 -- the 4-cell crab cavity as CSG: per half cell (cone ∩ tube) ∪ equator torus ∪
    (iris tube − iris torus), mirrored in z = 0, repeated along z and capped by a slab."""
    import math
    rho_sum = cav_rho + iris_rho
    rad_diff = cav_rad - iris_rad
    half2 = 0.25 * cell_len * cell_len
    diff2 = (rad_diff - rho_sum) ** 2
    cos_t = (rho_sum - rad_diff) * rho_sum
    cos_t += math.sqrt(half2 * (diff2 - rho_sum * rho_sum + half2))
    cos_t /= half2 + diff2
    theta = math.acos(cos_t)
    sin_t = math.sqrt(1 - cos_t * cos_t)
    cot_t = 1.0 / (sin_t / cos_t)
    cone_off = 0.5 * cell_len - iris_rho * sin_t + (iris_rad + iris_rho * (1.0 - cos_t)) * cot_t
    zhat, o = (0, 0, 1), (0, 0, 0)
    iris_tube = Shape.cylinder(iris_rad + iris_rho * (1.0 - cos_t), zhat, o)
    iris_torus = Shape.torus(iris_rad + iris_rho, iris_rho, zhat, (0, 0, 0.5 * cell_len))
    corr_iris_tube = Shape.subtract(iris_tube, iris_torus)
    cav_tube = Shape.cylinder(cav_rad - cav_rho * (1.0 - cos_t), zhat, o)
    cav_cone = Shape.cone(theta, zhat, (0, 0, cone_off))
    cav_torus = Shape.torus(cav_rad - cav_rho, cav_rho, zhat, o)
    pre_cav = Shape.intersection([cav_cone, cav_tube])
    half_cell = Shape.union([pre_cav, cav_torus, corr_iris_tube])
    full_cell = Shape.mirror(half_cell, zhat, o)
    inf_cells = Shape.repeat(full_cell, o, zhat, cell_len, num_cells // 2, num_cells // 2)
    caps = Shape.slab(float(num_cells) * cell_len, zhat, o)
    return Shape.intersection([caps, inf_cells])


def crabcav(cell_res=10, pad=2, num_cells=4, cell_len=2.0 * 0.0192, cav_rad=0.04719):
    """example/crabcav.py:69-91: grid of cell_res cells per cavity cell along z plus `pad` cells of metal around."""
    import math
    delta = cell_len / float(cell_res)
    nz = num_cells * cell_res + 2 * pad
    lz = float(nz) * delta
    nx = 2 * (int(math.ceil(cav_rad / delta)) + pad)
    lx = float(nx) * delta
    return Sim((nx, nx, nz), origin=(-0.5 * lx, -0.5 * lx, -0.5 * lz), size=(lx, lx, lz),
               pec=crabcav_shape(num_cells=num_cells, cell_len=cell_len, cav_rad=cav_rad))


def pill_w_tubes(cells_per_iris=4, ph_adv=2.0 * 3.141592653589793 / 3.0):
    """example/pillWTubes.py:14-150 -- one period of an iris-loaded 12 GHz accelerator structure: pillbox cavity
    (cylinder ∩ slab) united with the rounded iris tube (cylinder ∩ two inverted tori), periodic in z with the Bloch
    phase advance `ph_adv` -> complex operators on a Dey-Mittra PEC geometry. cells_per_iris = 4 is the example's grid."""
    import math
    c0, mm, ghz = 2.99792458e8, 1.0e-3, 1.0e9
    omega010 = 2.0 * math.pi * 12.0 * ghz
    iris_r, iris_t = 3.15 * mm, 1.67 * mm
    sync_ph_adv = 2.0 * math.pi / 3.0
    lz = c0 * sync_ph_adv / omega010
    lz_cav = lz - iris_t
    # getCavRadius (pillWTubes.py:22-62): first-order perturbation estimate of the iris detuning
    alpha = math.sqrt((2.405 / iris_r) ** 2 - (omega010 / c0) ** 2)
    tau = 2.0 / (math.pi * (2.405 * 0.5191) ** 2)
    t1 = tau * (iris_r ** 3 / (3.0 * lz_cav * c0 ** 2))
    t1 *= 1.0 - math.cos(sync_ph_adv) * math.exp(-alpha * iris_t)
    t2 = (math.sqrt(3.0 * 27.0 * t1 ** 4 * omega010 ** 2 + 12.0 * t1 ** 3) + 9.0 * t1 ** 2 * omega010) ** (1.0 / 3.0)
    omega_pill = t2 / (2.0 ** (1.0 / 3.0) * 3.0 ** (2.0 / 3.0) * t1) - (2.0 / 3.0) ** (1.0 / 3.0) / t2
    R = c0 * 2.405 / omega_pill
    zhat, zzz = (0, 0, 1), (0, 0, 0)
    cyl = Shape.cylinder(R, zhat, zzz)
    caps = Shape.slab(lz_cav, zhat, (0, 0, 0.5 * lz))
    pill = Shape.intersection([cyl, caps])
    iris_cyl = Shape.cylinder(iris_r + 0.5 * iris_t, zhat, zzz)
    torus1 = Shape.torus(iris_r + 0.5 * iris_t, 0.5 * iris_t, zhat, zzz).invert()
    torus2 = Shape.torus(iris_r + 0.5 * iris_t, 0.5 * iris_t, zhat, (0, 0, lz)).invert()
    iris_tube = Shape.intersection([iris_cyl, torus1, torus2])
    cav = Shape.union([pill, iris_tube])
    d = iris_t / float(cells_per_iris)
    lx = 2.0 * R + 4.0 * d
    nz = int(lz / d) + 1
    nx = int(lx / d) + 1
    sim = Sim((nx, nx, nz), origin=(-0.5 * lx, -0.5 * lx, 0.0), size=(lx, lx, lz), phase_shifts=(0.0, 0.0, ph_adv), pec=cav)
    sim.info = {"R": R, "lz": lz, "iris_r": iris_r, "iris_t": iris_t, "k2_target": (omega010 / c0) ** 2}
    return sim


def dsphmsph(n, eps=10.0, a=0.37, b=0.49, size=1.0, origin=-0.5):
    """example/dsphmsph.py: dielectric sphere (radius a, permittivity eps) inside a PEC sphere (radius b).
    The example runs one octant with PEC/PMC symmetry planes; this helper takes the full ball in a box."""
    diel = Shape.sphere(a, (0, 0, 0))
    metal = Shape.sphere(b, (0, 0, 0))
    e = np.eye(3) * eps if np.isscalar(eps) else np.asarray(eps)
    return Sim(n, origin=(origin,) * 3, size=(size,) * 3, pec=metal, dielectrics=[(diel, e)])


def oct_lower_bcs(pol, l, m, phase="re"):
    """example/dsphmsph.py:436-493 -- boundary conditions on the x = 0, y = 0, z = 0 symmetry planes that select the
    spherical multipole (pol 'TM' | 'TE', degree l, order m, phase 're' | 'im') when only one octant is simulated."""
    res = [PEC if m % 2 else PMC, PMC, PEC if (l - m) % 2 else PMC]
    flip = {PEC: PMC, PMC: PEC}
    if phase == "im":
        res[0], res[1] = flip[res[0]], flip[res[1]]
    elif phase != "re":
        raise ValueError("phase must be 're' or 'im'")
    if pol == "TE":
        res = [flip[b] for b in res]
    elif pol != "TM":
        raise ValueError("pol must be 'TM' or 'TE'")
    return tuple(res)


def dsphmsph_octant(n, lower, eps=10.0, a=0.37, b=0.49, size=0.5):
    """Config C3 as the reference runs it (example/run.py:14-29,169-179): the octant [0, 0.5]^3, PEC upper boundaries,
    PEC/PMC symmetry planes `lower` (see oct_lower_bcs), dielectric sphere inside a PEC sphere, both centred on the
    origin corner."""
    diel = Shape.sphere(a, (0, 0, 0))
    metal = Shape.sphere(b, (0, 0, 0))
    e = np.eye(3) * eps if np.isscalar(eps) else np.asarray(eps)
    return Sim(n, origin=(0.0,) * 3, size=(size,) * 3, lower=tuple(lower), upper=(PEC,) * 3, pec=metal, dielectrics=[(diel, e)])


# eps diag [10.225, 10.225, 9.95], off-diag [yz, xz, xy] = [0.6736.., -0.6736.., -0.825] (MxProblem.cpp:501-506)
_S = 0.67360967926537398
SAPPHIRE = np.array([[10.225, -0.825, -_S], [-0.825, 10.225, _S], [-_S, _S, 9.95]])


def phc_sapphire(n, phase_shifts=(0.0, 0.0, 0.0), r=0.37, eps=None):
    """example/phc-sapph-r0.37.py:12-22: sapphire sphere in a periodic unit cell, Bloch phase shifts."""
    sph = Shape.sphere(r, (0, 0, 0))
    return Sim(n, origin=(-0.5,) * 3, size=(1.0,) * 3, phase_shifts=phase_shifts,
               dielectrics=[(sph, SAPPHIRE if eps is None else eps)])


def vacuum(n, phase_shifts=None, literal=False):
    """example/vacuum.py:7-22 -- periodic unit box, no PEC."""
    return Sim(n, origin=(0.0,) * 3, size=(1.0,) * 3, phase_shifts=phase_shifts, literal_upper_periodic_e=literal)

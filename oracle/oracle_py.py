"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# Cowper 13-point rule (order 7): /root/reference/src/QuadratureFormula3d.cuh:181-213
QF13_XY = np.array([
    [0.333333333333333, 0.333333333333333], [0.479308067841923, 0.260345966079038],
    [0.260345966079038, 0.479308067841923], [0.260345966079038, 0.260345966079038],
    [0.869739794195598, 0.065130102902216], [0.065130102902216, 0.869739794195598],
    [0.065130102902216, 0.065130102902216], [0.638444188569809, 0.312865496004875],
    [0.312865496004875, 0.638444188569809], [0.638444188569809, 0.048690315425316],
    [0.048690315425316, 0.638444188569809], [0.312865496004875, 0.048690315425316],
    [0.048690315425316, 0.312865496004875]], dtype=np.float64)
QF13_W = np.array([-0.149570044467670, 0.175615257433204, 0.175615257433204, 0.175615257433204,
                   0.053347235608839, 0.053347235608839, 0.053347235608839, 0.077113760890257,
                   0.077113760890257, 0.077113760890257, 0.077113760890257, 0.077113760890257,
                   0.077113760890257], dtype=np.float64)
QF13_ORDER = 7

SIMPLE, ATTACHED, NOT = 0, 1, 2


def build() -> str:
    path = os.path.join(_HERE, "liboracle.so")
    subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"], check=True)
    return path


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "oracle.cpp")):
            build()
        L = C.CDLL(path)
        dp, ip, vp, ll = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_longlong
        L.orc_set_quadrature.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_mesh_create.restype = vp
        L.orc_mesh_create.argtypes = [dp, C.c_int, ip, C.c_int]
        L.orc_mesh_free.argtypes = [vp]
        L.orc_mesh_get.argtypes = [vp, dp, dp]
        L.orc_classify.argtypes = [vp, C.POINTER(ll)]
        L.orc_get_pairs.argtypes = [vp, C.c_int, ip]
        L.orc_theta_psi.argtypes = [dp, dp, dp, dp, dp]
        L.orc_singular_part.argtypes = [vp, C.c_int, dp, C.c_int, C.c_int, dp]
        L.orc_integrate_singular.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
        L.orc_regular_integrals.argtypes = [vp, C.c_int, ip, ll, C.c_int, dp]
        L.orc_regular_results_quad.argtypes = [vp, ip, ll, C.c_int, dp]
        L.orc_run_class.argtypes = [vp, C.c_int, ip, ll, C.c_int, dp, dp, C.POINTER(C.c_ubyte), C.POINTER(ll)]
        L.orc_run_class.restype = C.c_int
        L.orc_symmetry_error.argtypes = [dp, ll, dp]
        L.orc_noise_bound.argtypes = [vp, ip, ll, dp]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
        set_quadrature(QF13_XY, QF13_W, QF13_ORDER)
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def set_quadrature(xy, w, order):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    L = _LIB if _LIB is not None else lib()
    if L.orc_set_quadrature(_dp(xy), _dp(w), int(w.size), int(order)) != 0:
        raise ValueError("bad quadrature rule")


def num_threads() -> int:
    return int(lib().orc_num_threads())


def theta_psi(pt, A, B, Cc):
    out = np.empty(4)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pt, A, B, Cc)]
    lib().orc_theta_psi(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(out))
    return out


class OracleMesh:
    """Mesh + derived data as the reference computes them (normals, areas, neighbour lists)."""

    def __init__(self, vertices, cells):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self._h = lib().orc_mesh_create(_dp(self.vertices), self.vertices.shape[0], _ip(self.cells), self.cells.shape[0])
        self._pairs = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_mesh_free(self._h)
            self._h = None

    @property
    def n_cells(self):
        return self.cells.shape[0]

    def normals_measures(self):
        n = np.empty((self.n_cells, 3))
        s = np.empty(self.n_cells)
        lib().orc_mesh_get(self._h, _dp(n), _dp(s))
        return n, s

    def classify(self):
        """-> [simple, attached, not] lists of (i, j, k), i<j, sorted lexicographically."""
        if self._pairs is None:
            cnt = (C.c_longlong * 3)()
            lib().orc_classify(self._h, cnt)
            self._pairs = []
            for c in range(3):
                a = np.empty((cnt[c], 3), dtype=np.int32)
                if cnt[c]:
                    lib().orc_get_pairs(self._h, c, _ip(a))
                self._pairs.append(a)
        return self._pairs

    def tasks(self, cls):
        """Ordered task list of a class as Evaluator3D::runAllPairs builds it
        (/root/reference/src/evaluators/evaluator3d.cu:156-169): pairs then reversed pairs at n+idx."""
        p = self.classify()[cls]
        n = p.shape[0]
        t = np.empty((2 * n, 3), dtype=np.int32)
        t[:n] = p
        t[n:, 0] = p[:, 1]
        t[n:, 1] = p[:, 0]
        t[n:, 2] = n + np.arange(n, dtype=np.int32)
        return t

    def singular_part(self, cls, pt, i, j):
        out = np.empty(4)
        p = np.ascontiguousarray(pt, dtype=np.float64)
        lib().orc_singular_part(self._h, cls, _dp(p), int(i), int(j), _dp(out))
        return out

    def integrate_singular(self, cls, i, j):
        out = np.empty(4)
        lib().orc_integrate_singular(self._h, cls, int(i), int(j), _dp(out))
        return out

    def regular_integrals(self, cls, tasks, level=0):
        tasks = np.ascontiguousarray(tasks, dtype=np.int32)
        out = np.empty((tasks.shape[0], 4))
        lib().orc_regular_integrals(self._h, cls, _ip(tasks), tasks.shape[0], int(level), _dp(out))
        return out

    def regular_results_quad(self, tasks, level=0):
        """J of regular tasks evaluated with the reference's formulas in 113-bit arithmetic (the 'exact' value)."""
        tasks = np.ascontiguousarray(tasks, dtype=np.int32)
        out = np.empty((tasks.shape[0], 3))
        lib().orc_regular_results_quad(self._h, _ip(tasks), tasks.shape[0], int(level), _dp(out))
        return out

    def noise_bound(self, tasks):
        """first-order rounding-noise bound of the reference formula per regular task, level 0 (the tolerance model of
        tests/helpers.py::reference_noise_bound, evaluated in C/OpenMP for large task sets)"""
        tasks = np.ascontiguousarray(tasks, dtype=np.int32)
        out = np.empty(tasks.shape[0])
        lib().orc_noise_bound(self._h, _ip(tasks), tasks.shape[0], _dp(out))
        return out

    def run_class(self, cls, tasks, level=0):
        """Full path of one class. level<0 = adaptive. -> dict(integrals, results, refinements, stats, warn)"""
        tasks = np.ascontiguousarray(tasks, dtype=np.int32)
        n = tasks.shape[0]
        integrals = np.empty((n, 4))
        results = np.empty((n, 3))
        ref = np.zeros(self.n_cells, dtype=np.uint8)
        stats = (C.c_longlong * 13)()
        warn = lib().orc_run_class(self._h, cls, _ip(tasks), n, int(level), _dp(integrals), _dp(results),
                                   ref.ctypes.data_as(C.POINTER(C.c_ubyte)), stats)
        return dict(integrals=integrals, results=results, refinements=ref, stats=np.array(list(stats), dtype=np.int64), warn=warn)


def symmetry_error(results):
    results = np.ascontiguousarray(results, dtype=np.float64)
    n = results.shape[0] // 2
    err = np.empty(2 * n)
    lib().orc_symmetry_error(_dp(results), n, _dp(err))
    return err

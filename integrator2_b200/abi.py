"""ctypes binding of libintegrator2_b200.so (the C ABI in include/i2_abi.h).

This is plumbing for tests/ and bench.py: device memory comes from torch tensors (data_ptr()), streams from
torch.cuda.  The product is the shared library; there is no Python or CPU implementation of the path, and
every call fails loudly if the library is missing or no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libintegrator2_b200.so")

SIMPLE, ATTACHED, NOT = 0, 1, 2
LEVEL_ADAPTIVE = -1
MATH_STRICT, MATH_FAST, MATH_FAST_LIBDEVICE, MATH_FAST_POINTWISE = 0, 1, 2, 3
CLASS_NAMES = ("simple", "attached", "not")

# Cowper rules as shipped by the reference (/root/reference/src/QuadratureFormula3d.cuh:181-213): 13 points, order 7
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


class Stats(C.Structure):
    _fields_ = [("last_round", C.c_int), ("integrated", C.c_longlong * 6), ("unconverged", C.c_longlong * 6),
                ("orientation_warnings", C.c_int)]

    def as_dict(self):
        return dict(last_round=int(self.last_round), integrated=[int(x) for x in self.integrated],
                    unconverged=[int(x) for x in self.unconverged], orientation_warnings=int(self.orientation_warnings))


EXPORTS = [
    "i2_create", "i2_destroy", "i2_set_stream", "i2_synchronize", "i2_set_math_mode", "i2_error_string",
    "i2_set_quadrature", "i2_mesh_geometry", "i2_set_mesh", "i2_classify_count", "i2_classify_fill",
    "i2_add_reversed_pairs", "i2_integrate_class", "i2_integrate_pairs", "i2_integrate_all", "i2_symmetry_error", "i2_host_prepare", "i2_host_run",
    "i2_host_device_views", "i2_host_checksums", "i2_peer_alloc", "i2_peer_open", "i2_peer_close", "i2_peer_free", "i2_host_set_shard", "i2_host_shard", "i2_peak_rates", "i2_peak_dfma_three_operand", "i2_peak_dfma_with_integer", "i2_refine_mesh_once", "i2_launch_count", "i2_set_profiling", "i2_profile_last", "i2_selftest_math", "i2_apply_regular", "i2_apply_regular_adaptive",
    "i2_error_summary", "i2_host_reserve", "i2_mgpu_reserve", "i2_host_row_costs", "i2_host_run_rounds", "i2_host_last_rounds", "i2_host_refinements", "i2_host_run_finalize", "i2_host_fetch", "i2_mgpu_unique_id", "i2_mgpu_create_rank", "i2_mgpu_create_local", "i2_mgpu_destroy", "i2_mgpu_info", "i2_mgpu_context",
    "i2_mgpu_set_quadrature", "i2_mgpu_set_math_mode", "i2_mgpu_synchronize", "i2_mgpu_prepare", "i2_mgpu_shard", "i2_mgpu_set_results_target",
    "i2_mgpu_run", "i2_mgpu_checksums", "i2_mgpu_gather", "i2_mgpu_fetch", "i2_mgpu_refinements", "i2_mgpu_error_summary",
    "i2_mgpu_apply_prepare", "i2_mgpu_apply", "i2_mgpu_apply_result", "i2_apply_prepare", "i2_apply", "i2_apply_rounds", "i2_apply_last_rounds", "i2_apply_finish",
]

_lib = None


class I2Error(RuntimeError):
    pass


def load_library():
    """Load the CUDA extension; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise I2Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(make -C integrator2_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ll, i32 = C.c_void_p, C.c_longlong, C.c_int
    L.i2_create.argtypes = [C.POINTER(vp), i32]
    L.i2_destroy.argtypes = [vp]
    L.i2_set_stream.argtypes = [vp, vp]
    L.i2_synchronize.argtypes = [vp]
    L.i2_set_math_mode.argtypes = [vp, i32]
    L.i2_error_string.argtypes = [i32]
    L.i2_error_string.restype = C.c_char_p
    L.i2_set_quadrature.argtypes = [vp, vp, vp, i32, i32]
    L.i2_mesh_geometry.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp]
    L.i2_set_mesh.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.i2_classify_count.argtypes = [vp, vp, i32, C.POINTER(ll)]
    L.i2_classify_fill.argtypes = [vp, vp, i32, vp, vp, vp]
    L.i2_add_reversed_pairs.argtypes = [vp, vp, ll]
    L.i2_integrate_class.argtypes = [vp, i32, vp, ll, i32, vp, vp, vp, vp, C.POINTER(Stats)]
    L.i2_integrate_pairs.argtypes = [vp, i32, vp, ll, i32, vp, vp, vp, vp, C.POINTER(Stats)]
    L.i2_integrate_all.argtypes = [vp, C.POINTER(vp), C.POINTER(ll), i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(Stats)]
    L.i2_symmetry_error.argtypes = [vp, vp, ll, vp]
    L.i2_host_prepare.argtypes = [vp, vp, i32, vp, i32, C.POINTER(ll)]
    L.i2_host_run.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(Stats)]
    L.i2_host_checksums.argtypes = [vp, C.POINTER(C.c_double)]
    L.i2_host_device_views.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.i2_refine_mesh_once.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp, vp]
    L.i2_apply_regular.argtypes = [vp, i32, i32, vp, vp]
    L.i2_apply_regular_adaptive.argtypes = [vp, i32, i32, vp, vp, vp, vp, C.POINTER(Stats)]
    L.i2_selftest_math.argtypes = [vp, i32, vp, vp, ll, vp]
    L.i2_launch_count.argtypes = [C.POINTER(ll)]
    L.i2_set_profiling.argtypes = [vp, i32]
    L.i2_profile_last.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.i2_peak_dfma_three_operand.argtypes = [vp, C.POINTER(C.c_double)]
    L.i2_peak_dfma_with_integer.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.i2_peak_rates.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.i2_host_set_shard.argtypes = [vp, i32, i32]
    L.i2_host_shard.argtypes = [vp, C.POINTER(ll), C.POINTER(ll)]
    L.i2_peer_alloc.argtypes = [vp, C.c_ulonglong, C.POINTER(vp), C.c_char_p]
    L.i2_peer_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.i2_peer_close.argtypes = [vp, vp]
    L.i2_peer_free.argtypes = [vp, vp]
    L.i2_error_summary.argtypes = [vp, vp, ll, C.POINTER(C.c_double)]
    L.i2_host_reserve.argtypes = [vp, i32, i32]
    L.i2_mgpu_reserve.argtypes = [vp, i32, i32]
    L.i2_host_row_costs.argtypes = [vp, i32, vp, vp]
    L.i2_host_run_rounds.argtypes = [vp, i32]
    L.i2_host_last_rounds.argtypes = [vp, C.POINTER(i32), i32]
    L.i2_host_refinements.argtypes = [vp, vp, i32]
    L.i2_host_run_finalize.argtypes = [vp, i32, i32]
    L.i2_host_fetch.argtypes = [vp, i32, vp, vp, vp]
    L.i2_mgpu_unique_id.argtypes = [C.c_char_p]
    L.i2_mgpu_create_rank.argtypes = [C.POINTER(vp), i32, i32, i32, C.c_char_p]
    L.i2_mgpu_create_local.argtypes = [C.POINTER(vp), i32, C.POINTER(i32)]
    L.i2_mgpu_destroy.argtypes = [vp]
    L.i2_mgpu_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.i2_mgpu_context.argtypes = [vp, i32]
    L.i2_mgpu_context.restype = vp
    L.i2_mgpu_set_quadrature.argtypes = [vp, vp, vp, i32, i32]
    L.i2_mgpu_set_math_mode.argtypes = [vp, i32]
    L.i2_mgpu_synchronize.argtypes = [vp]
    L.i2_mgpu_prepare.argtypes = [vp, vp, i32, vp, i32, i32, C.POINTER(ll)]
    L.i2_mgpu_shard.argtypes = [vp, i32, C.POINTER(ll), C.POINTER(ll)]
    L.i2_mgpu_set_results_target.argtypes = [vp, i32, C.POINTER(vp)]
    L.i2_mgpu_run.argtypes = [vp, i32, i32, C.POINTER(Stats)]
    L.i2_mgpu_checksums.argtypes = [vp, C.POINTER(C.c_double)]
    L.i2_mgpu_gather.argtypes = [vp, i32, i32, i32, vp]
    L.i2_mgpu_fetch.argtypes = [vp, i32, i32, vp, vp, vp]
    L.i2_mgpu_refinements.argtypes = [vp, i32, vp]
    L.i2_mgpu_error_summary.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.i2_mgpu_apply_prepare.argtypes = [vp, vp, i32, vp, i32, i32, C.POINTER(i32)]
    L.i2_mgpu_apply.argtypes = [vp, i32, vp, vp, C.POINTER(Stats)]
    L.i2_mgpu_apply_result.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
    L.i2_apply_prepare.argtypes = [vp, i32, i32]
    L.i2_apply.argtypes = [vp, i32, vp, vp, vp, C.POINTER(Stats)]
    L.i2_apply_rounds.argtypes = [vp, i32, vp]
    L.i2_apply_last_rounds.argtypes = [vp, C.POINTER(i32), i32]
    L.i2_apply_finish.argtypes = [vp, i32, vp, vp, vp, C.POINTER(Stats)]
    for name in EXPORTS:
        if name not in ("i2_error_string", "i2_mgpu_context"):
            getattr(L, name).restype = i32
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise I2Error(f"i2 call failed ({rc}): {load_library().i2_error_string(rc).decode()}")


def _ptr(t):
    """device tensor, raw device address (int, e.g. a peer-mapped pointer from Context.peer_open) or None"""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class _RawCudaBuffer:
    """__cuda_array_interface__ view of a device allocation owned by the library (Context.peer_alloc)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def launch_count() -> int:
    n = C.c_longlong()
    _check(load_library().i2_launch_count(C.byref(n)))
    return int(n.value)


class Context:
    """One integrator context on one CUDA device (mirrors the lifetime of the reference's
    Mesh3D + NumericalIntegrator3D + EvaluatorJ3DK trio)."""

    def __init__(self, device: int = 0, math_mode: int = MATH_FAST, borrowed=None):
        import torch
        self.torch = torch
        self.L = load_library()
        self.device = device
        self.owned = borrowed is None
        if borrowed is None:
            h = C.c_void_p()
            _check(self.L.i2_create(C.byref(h), device))
        else:
            h = C.c_void_p(borrowed)        # a context owned by a MultiGpu handle
        self.h = h
        _check(self.L.i2_set_math_mode(self.h, math_mode))
        # run on torch's current stream: tensors handed to the library are produced/consumed by torch ops on that
        # stream, so kernels and torch copies stay ordered (a private stream would race with torch's fills/copies)
        if torch.cuda.is_available():
            with torch.cuda.device(device):
                self.set_stream(torch.cuda.current_stream().cuda_stream)
        self.set_quadrature(QF13_XY, QF13_W, QF13_ORDER)
        self._keep = []
        self.nc = 0

    def close(self):
        if getattr(self, "h", None):
            if self.owned:
                self.L.i2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def set_math_mode(self, mode):
        _check(self.L.i2_set_math_mode(self.h, mode))

    def set_quadrature(self, xy, w, order):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        _check(self.L.i2_set_quadrature(self.h, xy.ctypes.data, w.ctypes.data, int(w.size), int(order)))

    def synchronize(self):
        _check(self.L.i2_synchronize(self.h))

    # ---- device-pointer API ------------------------------------------------------------------------
    def set_mesh(self, vertices, cells):
        """vertices float64[nv,3] (scaled), cells int32[nc,3]; returns (normals, measures) device tensors."""
        torch = self.torch
        dev = f"cuda:{self.device}"
        v = torch.as_tensor(np.ascontiguousarray(vertices, dtype=np.float64)).to(dev)
        c = torch.as_tensor(np.ascontiguousarray(cells, dtype=np.int32)).to(dev)
        nv, nc = v.shape[0], c.shape[0]
        normals = torch.empty((nc, 3), dtype=torch.float64, device=dev)
        centers = torch.empty((nc, 3), dtype=torch.float64, device=dev)
        measures = torch.empty((nc,), dtype=torch.float64, device=dev)
        _check(self.L.i2_mesh_geometry(self.h, _ptr(v), nv, _ptr(c), nc, _ptr(normals), _ptr(centers), _ptr(measures)))
        _check(self.L.i2_set_mesh(self.h, _ptr(v), nv, _ptr(c), nc, _ptr(normals), _ptr(measures)))
        self._keep = [v, c, normals, measures, centers]
        self.d_vertices, self.d_cells, self.d_normals, self.d_measures, self.d_centers = v, c, normals, measures, centers
        self.nc = nc
        return normals, measures

    def classify(self, regular=True):
        """-> list of 3 int32[n,3] device tensors (i<j pairs, lexicographic, k = slot).  regular=False skips the regular
        list (None in its place; meshes whose N^2/2 list would not fit integrate that class list-free)."""
        torch = self.torch
        cnt = (C.c_longlong * 3)()
        _check(self.L.i2_classify_count(self.h, _ptr(self.d_cells), self.nc, cnt))
        self.pair_counts = [int(x) for x in cnt]
        dev = f"cuda:{self.device}"
        lists = [torch.empty((int(cnt[k]), 3), dtype=torch.int32, device=dev) if (k < 2 or regular) else None for k in range(3)]
        _check(self.L.i2_classify_fill(self.h, _ptr(self.d_cells), self.nc, _ptr(lists[0]), _ptr(lists[1]), _ptr(lists[2])))
        return lists

    def tasks_from_pairs(self, pairs):
        """pairs int32[n,3] device tensor -> ordered tasks int32[2n,3] (pairs + reversed pairs at n+idx)."""
        torch = self.torch
        n = pairs.shape[0]
        t = torch.empty((2 * n, 3), dtype=torch.int32, device=pairs.device)
        t[:n] = pairs
        _check(self.L.i2_add_reversed_pairs(self.h, _ptr(t), n))
        return t

    def integrate_pairs(self, cls, tasks, level=0, want_stats=True, refinements=None, out=None):
        """like integrate_class for a runAllPairs-shaped list [n/2 pairs ; n/2 reversed pairs] (i2_integrate_pairs)"""
        return self.integrate_class(cls, tasks, level, want_stats, refinements, out, pairs=True)

    def integrate_class(self, cls, tasks, level=0, want_stats=True, refinements=None, out=None, pairs=False):
        """tasks: int32[n,3] device tensor. -> dict(integrals[n,4], results[n,3], refinements, converged, stats)"""
        torch = self.torch
        n = tasks.shape[0]
        dev = tasks.device
        if out is None:
            integrals = torch.empty((n, 4), dtype=torch.float64, device=dev)
            results = torch.empty((n, 3), dtype=torch.float64, device=dev)
        else:
            integrals, results = out
        conv = None
        if level < 0:
            if refinements is None:
                refinements = torch.zeros((self.nc,), dtype=torch.uint8, device=dev)
            conv = torch.zeros((n,), dtype=torch.uint8, device=dev)
        st = Stats()
        fn, count = (self.L.i2_integrate_pairs, n // 2) if pairs else (self.L.i2_integrate_class, n)
        _check(fn(self.h, cls, _ptr(tasks), count, level, _ptr(integrals), _ptr(results),
                  _ptr(refinements) if level < 0 else C.c_void_p(0), _ptr(conv), C.byref(st) if want_stats else None))
        return dict(integrals=integrals, results=results, refinements=refinements, converged=conv,
                    stats=st.as_dict() if want_stats else None)

    def integrate_all(self, tasks, level=0, want_stats=True, refinements=None, out=None):
        """The three classes in one call (adjacent classes overlap the regular one on side streams).
        tasks: list of 3 int32[n_k,3] device tensors; out: optional list of 3 (integrals, results) pairs.
        -> list of 3 dicts like integrate_class."""
        torch = self.torch
        dev = tasks[0].device
        n = [int(t.shape[0]) for t in tasks]
        if out is None:
            out = [(torch.empty((n[k], 4), dtype=torch.float64, device=dev), torch.empty((n[k], 3), dtype=torch.float64, device=dev))
                   for k in range(3)]
        conv = [None] * 3
        if level < 0:
            if refinements is None:
                refinements = [torch.zeros((self.nc,), dtype=torch.uint8, device=dev) for _ in range(3)]
            conv = [torch.zeros((n[k],), dtype=torch.uint8, device=dev) for k in range(3)]
        else:
            refinements = [None] * 3

        def arr(ts):
            return (C.c_void_p * 3)(*[_ptr(t) for t in ts])
        st = (Stats * 3)()
        _check(self.L.i2_integrate_all(self.h, arr(tasks), (C.c_longlong * 3)(*n), level, arr([o[0] for o in out]), arr([o[1] for o in out]),
                                       arr(refinements), arr(conv), st if want_stats else None))
        return [dict(integrals=out[k][0], results=out[k][1], refinements=refinements[k], converged=conv[k],
                     stats=st[k].as_dict() if want_stats else None) for k in range(3)]

    def apply_regular(self, row_lo, row_hi, weights=None, out=None):
        """out[i-row_lo] = sum_{j not sharing a vertex with i} w_j J(K_i,K_j) without any task list (device tensors)."""
        torch = self.torch
        if out is None:
            out = torch.empty((row_hi - row_lo, 3), dtype=torch.float64, device=f"cuda:{self.device}")
        _check(self.L.i2_apply_regular(self.h, int(row_lo), int(row_hi), _ptr(weights), _ptr(out)))
        return out

    def apply_regular_adaptive(self, row_lo, row_hi, weights=None, want_stats=True):
        """Row sums of the regular class under automatic error control, list-free (see i2_apply_regular_adaptive).
        -> dict(out, other, refinements, stats)"""
        torch = self.torch
        dev = f"cuda:{self.device}"
        n = row_hi - row_lo
        out = torch.empty((n, 3), dtype=torch.float64, device=dev)
        other = torch.empty((n, 3), dtype=torch.float64, device=dev)
        ref = torch.zeros((n,), dtype=torch.uint8, device=dev)
        st = Stats()
        _check(self.L.i2_apply_regular_adaptive(self.h, int(row_lo), int(row_hi), _ptr(weights), _ptr(out), _ptr(other), _ptr(ref),
                                                C.byref(st) if want_stats else None))
        return dict(out=out, other=other, refinements=ref, stats=st.as_dict() if want_stats else None)

    def apply_prepare(self, row_lo, row_hi):
        _check(self.L.i2_apply_prepare(self.h, int(row_lo), int(row_hi)))
        self._apply_rows = (int(row_lo), int(row_hi))

    def apply(self, level, weights=None, want_stats=True, split=None):
        """The whole operator for the prepared row block: out[i] = sum_j w_j J(K_i,K_j) over all classes (i2_apply).
        split: callable(last_rounds) -> last_rounds, run between the two halves (what a multi-GPU caller all-reduces)."""
        torch = self.torch
        lo, hi = self._apply_rows
        dev = f"cuda:{self.device}"
        out = torch.empty((hi - lo, 3), dtype=torch.float64, device=dev)
        ref = torch.zeros((3, hi - lo), dtype=torch.uint8, device=dev)
        st = (Stats * 3)()
        if split is None:
            _check(self.L.i2_apply(self.h, int(level), _ptr(weights), _ptr(out), _ptr(ref), st if want_stats else None))
        else:
            _check(self.L.i2_apply_rounds(self.h, int(level), _ptr(weights)))
            a = (C.c_int * 3)()
            _check(self.L.i2_apply_last_rounds(self.h, a, 0))
            new = split([int(x) for x in a])
            _check(self.L.i2_apply_last_rounds(self.h, (C.c_int * 3)(*new), 1))
            _check(self.L.i2_apply_finish(self.h, int(level), _ptr(weights), _ptr(out), _ptr(ref), st if want_stats else None))
        return dict(out=out, refinements=ref, stats=[x.as_dict() for x in st] if want_stats else None)

    def symmetry_error(self, results):
        torch = self.torch
        n = results.shape[0]
        err = torch.empty((n,), dtype=torch.float64, device=results.device)
        _check(self.L.i2_symmetry_error(self.h, _ptr(results), n // 2, _ptr(err)))
        return err

    def error_summary(self, errors):
        """(max, mean) of a device tensor of (i,j)/(j,i) defects"""
        out = (C.c_double * 2)()
        _check(self.L.i2_error_summary(self.h, _ptr(errors), int(errors.numel()), out))
        return float(out[0]), float(out[1])

    def selftest_math(self, op, a, b=None):
        out = self.torch.empty_like(a)
        _check(self.L.i2_selftest_math(self.h, op, _ptr(a), _ptr(b), a.numel(), _ptr(out)))
        return out

    def set_profiling(self, on=True):
        _check(self.L.i2_set_profiling(self.h, 1 if on else 0))

    def profile_last(self):
        a, b = C.c_float(), C.c_float()
        _check(self.L.i2_profile_last(self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def set_stream(self, cuda_stream_ptr):
        _check(self.L.i2_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def peak_rates(self):
        a, b = C.c_double(), C.c_double()
        _check(self.L.i2_peak_rates(self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def peak_dfma_with_integer(self, n):
        a = C.c_double()
        _check(self.L.i2_peak_dfma_with_integer(self.h, int(n), C.byref(a)))
        return float(a.value)

    def peak_dfma_three_operand(self):
        a = C.c_double()
        _check(self.L.i2_peak_dfma_three_operand(self.h, C.byref(a)))
        return float(a.value)

    # ---- multi-GPU export over NVLink peer stores ------------------------------------------------------------
    def peer_alloc(self, rows, cols=3):
        """Owner side: float64[rows, cols] allocation of its own + the 64-byte CUDA IPC handle to ship to the other
        processes.  -> (tensor view on this device, handle bytes, raw address)"""
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        _check(self.L.i2_peer_alloc(self.h, int(max(1, rows)) * cols * 8, C.byref(p), h))
        view = self.torch.as_tensor(_RawCudaBuffer(p.value, (int(rows), cols), "<f8"), device=f"cuda:{self.device}")
        return view, h.raw, int(p.value)

    def peer_open(self, handle: bytes) -> int:
        """Writer side: map the owner's allocation into this process; returns the raw device address (valid on this
        context's device; pass `addr + 24 * first_slot` as the `results` of integrate_class / integrate_all)."""
        p = C.c_void_p()
        _check(self.L.i2_peer_open(self.h, C.create_string_buffer(handle, 64), C.byref(p)))
        return int(p.value)

    def peer_close(self, addr: int):
        _check(self.L.i2_peer_close(self.h, C.c_void_p(addr)))

    def peer_free(self, addr: int):
        _check(self.L.i2_peer_free(self.h, C.c_void_p(addr)))

    # ---- host-buffer API (end-to-end) ------------------------------------------------------------------
    def host_prepare(self, vertices, cells):
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = np.ascontiguousarray(cells, dtype=np.int32)
        cnt = (C.c_longlong * 3)()
        _check(self.L.i2_host_prepare(self.h, v.ctypes.data, v.shape[0], c.ctypes.data, c.shape[0], cnt))
        self.nc = c.shape[0]
        return [int(x) for x in cnt]

    def host_set_shard(self, rank, world):
        """multi-GPU: this context integrates shard `rank` of `world` of every class (call before host_prepare)"""
        _check(self.L.i2_host_set_shard(self.h, int(rank), int(world)))

    def host_shard(self):
        a, b = (C.c_longlong * 3)(), (C.c_longlong * 3)()
        _check(self.L.i2_host_shard(self.h, a, b))
        return [int(x) for x in a], [int(x) for x in b]

    def host_run(self, level, h_tasks, h_results, h_errors=None, h_refinements=None):
        """h_* : lists of 3 pinned CPU torch tensors (or None entries). Returns list of 3 stats dicts."""
        def arr(lst):
            a = (C.c_void_p * 3)()
            for k in range(3):
                t = lst[k] if lst is not None else None
                a[k] = t.data_ptr() if t is not None and t.numel() else None
            return a
        st = (Stats * 3)()
        _check(self.L.i2_host_run(self.h, level, arr(h_tasks), arr(h_results), arr(h_errors), arr(h_refinements), st))
        return [s.as_dict() for s in st]

    def host_run_rounds(self, level):
        _check(self.L.i2_host_run_rounds(self.h, int(level)))

    def host_last_rounds(self, set_to=None):
        a = (C.c_int * 3)(*(set_to if set_to is not None else (0, 0, 0)))
        _check(self.L.i2_host_last_rounds(self.h, a, 1 if set_to is not None else 0))
        return [int(x) for x in a]

    def host_refinements(self, set_to=None):
        r = np.ascontiguousarray(set_to, dtype=np.uint8) if set_to is not None else np.empty((3, self.nc), dtype=np.uint8)
        _check(self.L.i2_host_refinements(self.h, r.ctypes.data, 1 if set_to is not None else 0))
        return r

    def host_run_finalize(self, level, check=False):
        _check(self.L.i2_host_run_finalize(self.h, int(level), 1 if check else 0))

    def host_fetch(self, cls, tasks=True, results=True, errors=False):
        n = self.host_shard()[1][cls]
        t = np.empty((n, 3), dtype=np.int32) if tasks else None
        r = np.empty((n, 3), dtype=np.float64) if results else None
        e = np.empty((n,), dtype=np.float64) if errors else None
        _check(self.L.i2_host_fetch(self.h, int(cls), t.ctypes.data if tasks else None, r.ctypes.data if results else None,
                                    e.ctypes.data if errors else None))
        return dict(tasks=t, results=r, errors=e)

    def host_row_costs(self, upper_only=True):
        """(predicted adaptive cost per row, first regular forward slot of every row [nc+1]) — see i2_host_row_costs"""
        cost = np.empty(self.nc, dtype=np.float64)
        first = np.empty(self.nc + 1, dtype=np.uint64)
        _check(self.L.i2_host_row_costs(self.h, 1 if upper_only else 0, cost.ctypes.data, first.ctypes.data))
        return cost, first

    def host_checksums(self):
        a = (C.c_double * 12)()
        _check(self.L.i2_host_checksums(self.h, a))
        return np.array(list(a)).reshape(3, 4)

    def host_device_views(self):
        t = (C.c_void_p * 3)()
        r = (C.c_void_p * 3)()
        _check(self.L.i2_host_device_views(self.h, t, r))
        return list(t), list(r)


class MultiGpu:
    """The multi-GPU layer of the C ABI (i2_mgpu_*): the task lists of runAllPairs sharded over the GPUs of one box.
    MultiGpu(local_gpus=N)                       one process drives N GPUs (the CLI / host-class mode)
    MultiGpu(device=d, rank=r, world=w, uid=..)  one process per GPU (torchrun); uid = MultiGpu.unique_id() of rank 0"""

    def __init__(self, local_gpus=None, device=0, rank=0, world=1, uid=None, math_mode=MATH_FAST):
        self.L = load_library()
        h = C.c_void_p()
        if local_gpus is not None:
            _check(self.L.i2_mgpu_create_local(C.byref(h), int(local_gpus), None))
        else:
            _check(self.L.i2_mgpu_create_rank(C.byref(h), int(device), int(rank), int(world), uid))
        self.h = h
        w, nl, fr = C.c_int(), C.c_int(), C.c_int()
        _check(self.L.i2_mgpu_info(self.h, C.byref(w), C.byref(nl), C.byref(fr)))
        self.world, self.n_local, self.first_rank = int(w.value), int(nl.value), int(fr.value)
        devs = list(range(self.n_local)) if local_gpus is not None else [int(device)]
        self.contexts = [Context(devs[k], math_mode, borrowed=self.L.i2_mgpu_context(self.h, k)) for k in range(self.n_local)]
        self.nc = 0

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(load_library().i2_mgpu_unique_id(buf))
        return buf.raw

    def close(self):
        if getattr(self, "h", None):
            for c in self.contexts:
                c.close()
            self.L.i2_mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def prepare(self, vertices, cells, level=0):
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = np.ascontiguousarray(cells, dtype=np.int32)
        cnt = (C.c_longlong * 3)()
        _check(self.L.i2_mgpu_prepare(self.h, v.ctypes.data, v.shape[0], c.ctypes.data, c.shape[0], int(level), cnt))
        self.nc = c.shape[0]
        for ctx in self.contexts:
            ctx.nc = self.nc
        return [int(x) for x in cnt]

    def shard(self, rank):
        a, b = (C.c_longlong * 3)(), (C.c_longlong * 3)()
        _check(self.L.i2_mgpu_shard(self.h, int(rank), a, b))
        return [int(x) for x in a], [int(x) for x in b]

    def run(self, level, check=False, want_stats=False):
        st = (Stats * 3)()
        _check(self.L.i2_mgpu_run(self.h, int(level), 1 if check else 0, st if want_stats else None))
        return [s.as_dict() for s in st] if want_stats else None

    def synchronize(self):
        _check(self.L.i2_mgpu_synchronize(self.h))

    def checksums(self):
        a = (C.c_double * 12)()
        _check(self.L.i2_mgpu_checksums(self.h, a))
        return np.array(list(a)).reshape(3, 4)

    def fetch(self, local_index, cls, tasks=True, results=True, errors=False):
        """row-striped export of one local GPU's shard -> dict of numpy arrays in the shard's own order"""
        n = self.shard(self.first_rank + local_index)[1][cls]
        out = {}
        t = np.empty((n, 3), dtype=np.int32) if tasks else None
        r = np.empty((n, 3), dtype=np.float64) if results else None
        e = np.empty((n,), dtype=np.float64) if errors else None
        _check(self.L.i2_mgpu_fetch(self.h, int(local_index), int(cls), t.ctypes.data if tasks else None, r.ctypes.data if results else None,
                                    e.ctypes.data if errors else None))
        out.update(tasks=t, results=r, errors=e)
        return out

    def gather(self, cls, what, root, dst):
        """result (what=0, float64[n,3]) or task (what=1, int32[n,3]) shards of a class concatenated in rank order into the device
        tensor `dst` on GPU `root` (None on processes that do not own the root)"""
        _check(self.L.i2_mgpu_gather(self.h, int(cls), int(what), int(root), _ptr(dst)))

    def error_summary(self, cls):
        out = (C.c_double * 2)()
        _check(self.L.i2_mgpu_error_summary(self.h, int(cls), out))
        return float(out[0]), float(out[1])

    def set_results_target(self, local_index, ptrs):
        arr = (C.c_void_p * 3)(*[_ptr(p) for p in ptrs]) if ptrs is not None else None
        _check(self.L.i2_mgpu_set_results_target(self.h, int(local_index), arr))

    def apply_prepare(self, vertices, cells, level=0):
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = np.ascontiguousarray(cells, dtype=np.int32)
        cuts = (C.c_int * (self.world + 1))()
        _check(self.L.i2_mgpu_apply_prepare(self.h, v.ctypes.data, v.shape[0], c.ctypes.data, c.shape[0], int(level), cuts))
        self.nc = c.shape[0]
        for ctx in self.contexts:
            ctx.nc = self.nc
        self.row_cuts = [int(x) for x in cuts]
        return self.row_cuts

    def apply(self, level, weights=None, want_out=True, want_stats=False):
        w = np.ascontiguousarray(weights, dtype=np.float64) if weights is not None else None
        out = np.empty((self.nc, 3), dtype=np.float64) if want_out else None
        st = (Stats * 3)()
        _check(self.L.i2_mgpu_apply(self.h, int(level), w.ctypes.data if w is not None else None, out.ctypes.data if want_out else None,
                                    st if want_stats else None))
        return out, ([x.as_dict() for x in st] if want_stats else None)

    def apply_result(self, local_index=0):
        """(device tensor view float64[nc,3] of the full vector on local GPU k, uint8[3, rows of its block])"""
        import torch
        a, b = C.c_void_p(), C.c_void_p()
        _check(self.L.i2_mgpu_apply_result(self.h, int(local_index), C.byref(a), C.byref(b)))
        g = self.first_rank + local_index
        rows = self.row_cuts[g + 1] - self.row_cuts[g]
        dev = f"cuda:{self.contexts[local_index].device}"
        full = torch.as_tensor(_RawCudaBuffer(a.value, (self.nc, 3), "<f8"), device=dev)
        ref = torch.as_tensor(_RawCudaBuffer(b.value, (3, max(rows, 1)), "|u1"), device=dev)[:, :rows]
        return full, ref

    def refinements(self, cls):
        r = np.empty((self.nc,), dtype=np.uint8)
        _check(self.L.i2_mgpu_refinements(self.h, int(cls), r.ctypes.data))
        return r

"""CPU pre-check of the product's regular-pair arithmetic: tests/host_emu compiles the SAME device headers
(integrator2_b200/csrc/i2_pair.cuh, i2_math.cuh) with g++ and walks them in the order of k_regular_grouped (projection
form, symmetric log arguments, far-field shortcuts, careful-form redo), with the lane's own predicate in place of the warp
votes.  It is test infrastructure (never loaded by the product) and lets the numerics of the default kernel variant be
checked here, without a GPU, against the reference's golden results, the oracle and the 113-bit exact evaluation."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import check_parity_perturbation, check_regular_parity, couple_tasks, near_singular_mesh
from integrator2_b200.meshio import load_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_b200.npz"))
MODE_DEFAULT = 3 | (15 << 3)     # grouped, EDGELEN + no residual + DERIVE + projection form = kernel variant 27
RFP = 0.079577471545947667884
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def emu(oracle):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "host_emu")], check=True)
    lib = C.CDLL(os.path.join(ROOT, "tests", "host_emu", "libemu.so"))
    lib.emu_set_quadrature(oracle.QF13_XY.ctypes.data_as(dp), oracle.QF13_W.ctypes.data_as(dp), 13)
    return lib


def _emu_J(emu, om, tasks, level=0, mode=MODE_DEFAULT):
    nrm, S = om.normals_measures()
    t = np.ascontiguousarray(tasks, dtype=np.int32)
    o = np.empty((t.shape[0], 4))
    if level == 0:
        emu.emu_regular(om.vertices.ctypes.data_as(dp), om.cells.ctypes.data_as(ip), S.ctypes.data_as(dp), t.ctypes.data_as(ip),
                        C.c_longlong(t.shape[0]), mode, o.ctypes.data_as(dp))
    else:
        emu.emu_regular_level(om.vertices.ctypes.data_as(dp), om.cells.ctypes.data_as(ip), S.ctypes.data_as(dp), t.ctypes.data_as(ip),
                              C.c_longlong(t.shape[0]), mode, level, o.ctypes.data_as(dp))
    n = nrm[t[:, 1]]
    return RFP * (o[:, 3:4] * n + np.cross(o[:, :3], n))


@pytest.mark.parametrize("name,fixture,scale", [("G1_r0", "G1", 1.0), ("cubehole_r0", "cubehole", 1.0), ("s5m_r0", "s5m", 0.0005)])
def test_default_variant_matches_reference_golden(emu, oracle, name, fixture, scale):
    m = load_fixture(fixture, scale)
    om = oracle.OracleMesh(m.vertices, m.cells)
    t = np.ascontiguousarray(G[name + ".not.tasks"])
    J = _emu_J(emu, om, t)
    check_regular_parity(m.vertices, m.cells, t, J, G[name + ".not.J"], label=name)


@pytest.mark.parametrize("name,fixture,scale", [("G1_r0", "G1", 1.0), ("cubehole_r0", "cubehole", 1.0), ("s5m_r0", "s5m", 0.0005)])
def test_default_variant_is_not_noisier_than_the_reference(emu, oracle, name, fixture, scale):
    """distance to the exact value of the formulas (113-bit evaluation in the oracle), against the reference's own"""
    m = load_fixture(fixture, scale)
    om = oracle.OracleMesh(m.vertices, m.cells)
    t = np.ascontiguousarray(G[name + ".not.tasks"])[:200000]
    J = _emu_J(emu, om, t)
    exact = om.regular_results_quad(t, 0)
    scale_ = np.abs(exact).sum(1)
    e_new = np.abs(J - exact).sum(1) / scale_
    e_ref = np.abs(G[name + ".not.J"][:200000] - exact).sum(1) / scale_
    assert np.median(e_new) <= 1.5 * np.median(e_ref) + 1e-15
    assert np.quantile(e_new, 0.99) <= 3.0 * np.quantile(e_ref, 0.99) + 1e-14
    assert e_new.max() <= 5.0 * e_ref.max() + 1e-13


@pytest.mark.parametrize("level", [0, 1])
def test_nearly_singular_regular_pairs(emu, oracle, level):
    """Gauss points next to an edge / a vertex / on an edge line of the influence triangle: the careful-form redo
    (near_edge, near_vertex, large solid angle) and the reference's epsilon fallback."""
    v, c, couples, labels = near_singular_mesh()
    om = oracle.OracleMesh(v, c)
    t = couple_tasks(om, couples)
    assert t.shape[0] == 2 * couples.shape[0]
    J = _emu_J(emu, om, t, level)
    assert np.isfinite(J).all()
    check_parity_perturbation(oracle, v, c, 2, t, level, J, label=f"near-singular level {level}")

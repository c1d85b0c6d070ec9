"""CPU tests of the oracle (oracle/oracle.cpp): analytic known-answer checks of SURVEY.md §8(c) and the behaviour of
the adaptive loop.  The pinning of the oracle against dumps of the reference's own CUDA build is in
test_golden_reference.py."""
import numpy as np
import pytest

from integrator2_b200.meshio import TriMesh, fixture_names, load_fixture, parse_dat, subdivide

TWO_TRIANGLE = [n for n in [
    "Case-1-1", "Case-1-2", "Case-2-3", "Case-3-4", "Case-4-1", "Case-5-2", "Case-6-4", "Case-7-1", "Case-7-4", "Case-8-2",
    "Case-8-3", "Case-9-1", "Case1", "Case1_2", "Case1_vertex", "G1Sosed", "G1new", "G1Cont", "G1contactR", "genCase", "Test", "Case1-0"]]


def test_fixture_counts_match_the_survey():
    # loaded triangle counts: headers count line elements too
    for name, nc in (("G1", 106), ("s5m", 1864), ("s5m2", 7830), ("Vint16k", 16930)):
        assert load_fixture(name).n_cells == nc
    assert len([n for n in fixture_names() if n.startswith("Case-")]) == 33


def test_dat_parser_skips_line_elements():
    txt = "4 3\n1 0 0 0\n2 1 0 0\n3 0 1 0\n4 0 0 1\n1 102 1 2\n2 203 1 2 3\n3 203 1 3 4\n"
    m = parse_dat(txt, 2.0)
    assert m.n_cells == 2 and m.vertices[1, 0] == 2.0
    assert m.cells.tolist() == [[0, 1, 2], [0, 2, 3]]


def test_classification_counts_G1(oracle):
    m = load_fixture("G1")
    om = oracle.OracleMesh(m.vertices, m.cells)
    p = om.classify()
    assert [x.shape[0] for x in p] == [449, 159, 4957]          # ordered tasks 898 / 318 / 9914
    assert sum(x.shape[0] for x in p) == 106 * 105 // 2           # no pair shares 3 vertices


def test_antisymmetry_G1(oracle):
    """J(K_i,K_j) = -J(K_j,K_i): the reference's own --checkresults oracle (evaluator3d.cu:45-57)."""
    m = load_fixture("G1")
    om = oracle.OracleMesh(m.vertices, m.cells)
    expect = {0: (7.2e-7, 8.6e-6), 1: (1.5e-7, 2.3e-6), 2: (9.2e-9, 2.1e-5)}   # SURVEY.md §4 table (median, max)
    for cls in range(3):
        r = om.run_class(cls, om.tasks(cls), 0)
        err = oracle.symmetry_error(r["results"])
        assert np.isfinite(r["results"]).all()
        assert np.median(err) < expect[cls][0] * 1.05 and err.max() < expect[cls][1] * 1.05


def test_scale_and_translation_invariance(oracle):
    """J scales with the square of the mesh scale and is translation invariant."""
    m = load_fixture("G1")
    base = oracle.OracleMesh(m.vertices, m.cells)
    t = base.tasks(2)[::7]
    J0 = base.run_class(2, t, 0)["results"]
    s = 0.37
    J1 = oracle.OracleMesh(m.vertices * s, m.cells).run_class(2, t, 0)["results"]
    assert np.abs(J1 - s * s * J0).sum(1).max() <= 1e-12 * np.abs(s * s * J0).sum(1).max()
    J2 = oracle.OracleMesh(m.vertices + np.array([3.0, -2.0, 5.0]), m.cells).run_class(2, t, 0)["results"]
    assert (np.abs(J2 - J0).sum(1) / np.abs(J0).sum(1)).max() < 1e-9


def test_refinement_converges(oracle):
    """fixed levels converge: |I_2 - I_1| << |I_1 - I_0| for regular pairs (7th-order rule)."""
    m = load_fixture("G1")
    om = oracle.OracleMesh(m.vertices, m.cells)
    t = om.tasks(2)[::11]
    I0, I1, I2 = (om.regular_integrals(2, t, l) for l in (0, 1, 2))
    d01 = np.abs(I1 - I0).sum(1)
    d12 = np.abs(I2 - I1).sum(1)
    assert np.median(d12 / np.maximum(d01, 1e-300)) < 0.05


def test_adaptive_loop_G1(oracle):
    """Adaptive profile of the regular class on G1 (SURVEY.md §7 prediction) incl. the buffer ping-pong (D7)."""
    m = load_fixture("G1")
    om = oracle.OracleMesh(m.vertices, m.cells)
    t = om.tasks(2)
    r = om.run_class(2, t, -1)
    st = r["stats"]
    assert st[0] == 2                                  # L = 2 rounds
    assert st[1] == 9914 and st[3] == 39656 and st[4] == 5 and st[5] == 80 and st[6] == 0
    assert np.bincount(r["refinements"]).tolist() == [0, 101, 5]
    # D7: L is even -> tasks that converged in round 1 end with their UNREFINED value
    fixed0 = om.run_class(2, t, 0)["integrals"]
    same = np.all(r["integrals"] == fixed0, axis=1)
    assert same.sum() == 9914 - 5


@pytest.mark.parametrize("name", TWO_TRIANGLE)
def test_two_triangle_fixtures_run(oracle, name):
    """The special-case branches of the analytic singular integrals produce finite values."""
    m = load_fixture(name)
    om = oracle.OracleMesh(m.vertices, m.cells)
    p = om.classify()
    assert sum(x.shape[0] for x in p) == 1
    cls = [k for k in range(3) if p[k].shape[0]][0]
    r = om.run_class(cls, om.tasks(cls), 0)
    assert np.isfinite(r["results"]).all(), name
    a = om.run_class(cls, om.tasks(cls), -1)
    assert np.isfinite(a["results"]).all(), name
    assert 1 <= a["stats"][0] <= 5


def test_subdivide_matches_child_enumeration(oracle):
    """Integrating over the 4 children of a panel at level 0 equals integrating the panel at level 1."""
    m = load_fixture("G1")
    sub = subdivide(m, 1)
    assert sub.n_cells == 4 * m.n_cells
    om = oracle.OracleMesh(m.vertices, m.cells)
    t = om.tasks(2)[:40]
    I1 = om.regular_integrals(2, t, 1)
    # children of cell i are cells 4i..4i+3 of the subdivided mesh; influence panel j stays ORIGINAL -> build a mixed mesh
    verts = np.vstack([sub.vertices, m.vertices])
    off = sub.vertices.shape[0]
    cells = np.vstack([sub.cells, m.cells + off])
    mixed = oracle.OracleMesh(verts, cells)
    acc = np.zeros((t.shape[0], 4))
    for k in range(4):
        tt = np.stack([4 * t[:, 0] + k, t[:, 1] + sub.n_cells, t[:, 2]], axis=1).astype(np.int32)
        acc += mixed.regular_integrals(2, tt, 0)
    assert np.abs(acc - I1).max() <= 1e-13 * np.abs(I1).max()


def test_noise_bound_c_twin_matches_the_numpy_model(oracle):
    """orc_noise_bound (C/OpenMP, used for the 1e7-pair blocks of the largest-mesh parity tests) == helpers.reference_noise_bound"""
    from helpers import reference_noise_bound
    from integrator2_b200.meshio import load_fixture
    m = load_fixture("s5m", 0.0005)
    om = oracle.OracleMesh(m.vertices, m.cells)
    t = om.tasks(2)[::97]
    a, b = reference_noise_bound(m.vertices, m.cells, t), om.noise_bound(t)
    assert np.allclose(a, b, rtol=1e-6, atol=0)

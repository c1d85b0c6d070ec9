"""The drop-in C++ host stack end to end: tests/compat/ref_compat_dump is the SAME harness source that ran against the
reference (oracle/ref_dump.cu, written against the reference's public headers), compiled against include/integrator2/.
Its dumps are compared with the reference's dumps (tests/golden/reference_b200.npz) keyed on (i, j): this exercises
Mesh3D::loadMeshFromFile / prepareMesh, NumericalIntegrator3D, Evaluator3D::runAllPairs / runPairs and EvaluatorJ3DK."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from helpers import align_by_pair, check_parity_perturbation, read_class_dump, read_mesh_dump
from integrator2_b200.meshio import load_fixture, write_dat
from test_golden_reference import CLS, G, META, parse_rounds

pytestmark = pytest.mark.gpu
DUMP = os.path.join(ROOT, "tests", "compat", "ref_compat_dump")


def _dump(tmp_path, mesh_name, scale, args):
    m = load_fixture(mesh_name)                       # unscaled: the loader applies -s itself, like the reference
    write_dat(str(tmp_path / "mesh.dat"), m)
    cmd = [DUMP, "-f", "mesh.dat", "-o", "out"] + (["-s", repr(scale)] if scale != 1.0 else []) + args
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("name,mesh,scale,args", [("G1_r0", "G1", 1.0, ["-r", "0"]), ("G1_r1", "G1", 1.0, ["-r", "1"]), ("G1_ad", "G1", 1.0, [])])
def test_dropin_classes_match_reference_dumps_G1(tmp_path, name, mesh, scale, args):
    out = _dump(tmp_path, mesh, scale, args)
    md = read_mesh_dump(str(tmp_path / "out.mesh.bin"))
    ref_mesh = load_fixture(mesh, scale)
    assert np.array_equal(md["cells"], ref_mesh.cells) and np.array_equal(md["vertices"], ref_mesh.vertices)
    for c, cn in enumerate(CLS):
        d = read_class_dump(str(tmp_path / f"out.{cn}.bin"))
        ia, ib = align_by_pair(d["tasks"], G[f"{name}.{cn}.tasks"])
        assert ia.size == d["tasks"].shape[0] == G[f"{name}.{cn}.tasks"].shape[0]
        J, Jr = d["results"][ia], G[f"{name}.{cn}.J"][ib]
        assert (np.abs(J - Jr).sum(1) / np.abs(Jr).sum(1)).max() <= 1e-12, (name, cn)
        if f"{name}.{cn}.I" in G.files:
            I, Ir = d["integrals"][ia], G[f"{name}.{cn}.I"][ib]
            assert (np.abs(I - Ir).sum(1) / np.abs(Ir).sum(1)).max() <= 1e-12
    if not args:
        for c in range(3):
            assert np.array_equal(md["refinements"][c], G[f"{name}.refinements"][c])
        mine = [ln for ln in out.splitlines() if ln.startswith(("Iteration", "Out of"))]
        ref = [ln for ln in META[name]["log"] if ln.startswith(("Iteration", "Out of"))]
        assert mine == ref


def test_dropin_scaled_airplane_matches_reference_dumps(tmp_path, oracle):
    """s5m.dat with -s 0.0005 (the README example of the reference), fixed level 0."""
    _dump(tmp_path, "s5m", 0.0005, ["-r", "0"])
    m = load_fixture("s5m", 0.0005)
    for c, cn in enumerate(CLS[:2]):
        d = read_class_dump(str(tmp_path / f"out.{cn}.bin"))
        ia, ib = align_by_pair(d["tasks"], G[f"s5m_r0.{cn}.tasks"])
        assert ib.size == G[f"s5m_r0.{cn}.tasks"].shape[0]
        check_parity_perturbation(oracle, m.vertices, m.cells, c, d["tasks"][ia], 0, d["results"][ia], J_ref=G[f"s5m_r0.{cn}.J"][ib], label=cn)


def test_dropin_run_pairs_user_lists(tmp_path, oracle):
    """Evaluator3D::runPairs with user-supplied task lists (src/evaluators/evaluator3d.cu:213-288)."""
    m = load_fixture("G1")
    om = oracle.OracleMesh(m.vertices, m.cells)
    lists = [np.ascontiguousarray(om.tasks(c)[::3]) for c in range(3)]
    for L in lists:
        L[:, 2] = np.arange(L.shape[0])               # the user must number its own list (reference convention)
    with open(tmp_path / "pairs.bin", "wb") as f:
        f.write(np.array([L.shape[0] for L in lists], dtype=np.int32).tobytes())
        for L in lists:
            f.write(L.astype(np.int32).tobytes())
    _dump(tmp_path, "G1", 1.0, ["-r", "1", "--pairs", "pairs.bin"])
    for c, cn in enumerate(CLS):
        d = read_class_dump(str(tmp_path / f"out.{cn}.bin"))
        assert np.array_equal(d["tasks"], lists[c])
        ref = om.run_class(c, lists[c], 1)
        assert (np.abs(d["results"] - ref["results"]).sum(1) / np.abs(ref["results"]).sum(1)).max() <= 1e-12

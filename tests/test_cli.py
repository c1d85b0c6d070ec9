"""Drop-in CLI `integrator2test3D` (integrator2_b200/host/cli_main.cpp): flags, stdout lines and export formats of the
reference's CLI (/root/reference/tests/integrator3D/main.cu, README flag table)."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from integrator2_b200.meshio import load_fixture, write_dat

CLI = os.path.join(ROOT, "integrator2_b200", "host", "integrator2test3D")


def _run(args, cwd):
    return subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=600)


def test_cli_argument_handling_without_gpu(tmp_path):
    """Usage / argument errors are decided before any CUDA call (same exit codes as the reference)."""
    assert os.path.exists(CLI), "build with make -C integrator2_b200/host"
    r = _run([], tmp_path)
    assert r.returncode == 0 and "USAGE: integrator2test3D [options]" in r.stdout and "--exporttocsv" in r.stdout
    assert _run(["--help"], tmp_path).returncode == 0
    r = _run(["-c"], tmp_path)
    assert r.returncode != 0 and "No input file with mesh specified. Exiting" in r.stdout
    r = _run(["--bogus"], tmp_path)
    assert r.returncode != 0 and "Unknown option '--bogus'" in r.stderr
    r = _run(["-f", "x.dat", "-s", "abc"], tmp_path)
    assert r.returncode != 0 and "requires a numeric argument" in r.stderr


@pytest.mark.gpu
def test_cli_fixed_level_csv_matches_oracle(tmp_path, oracle):
    m = load_fixture("G1")
    write_dat(str(tmp_path / "G1.dat"), m)
    r = _run(["-f", "G1.dat", "-r", "0", "-c", "--exporttocsv", "--exporttoobj", "--exporttovtk"], tmp_path)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "Loaded mesh with 55 vertices and 106 cells" in out
    assert "Found 449 pairs of simple neighbors and 159 pairs of attached neighbors, 4957 pairs are not neighbors" in out
    assert "Using original mesh without refinement" in out
    for cls_line in ("Integrating over simple neighbors (898 pairs)...", "Integrating over attached neighbors (318 pairs)...",
                     "Integrating over not neighbors (9914 pairs)..."):
        assert cls_line in out
    assert len(re.findall(r"Time for (Simple neighbors|Attached neighbors|Non-neighbors) integration:\s+[0-9.]+ ms", out)) == 3
    assert "9914 results saved to file NotNeighbors.csv" in out
    om = oracle.OracleMesh(m.vertices, m.cells)
    for cls, fn in enumerate(("SimpleNeighbors.csv", "AttachedNeighbors.csv", "NotNeighbors.csv")):
        lines = open(tmp_path / fn).read().splitlines()
        assert lines[0] == '"TaskI";"TaskJ";"IntegralX";"IntegralY";"IntegralZ";"Error"'
        rows = np.array([[float(x) for x in ln.split(";")] for ln in lines[1:]])
        tasks = om.tasks(cls)
        assert np.array_equal(rows[:, :2].astype(np.int32), tasks[:, :2])
        ref = om.run_class(cls, tasks, 0)
        assert np.allclose(rows[:, 2:5], ref["results"], rtol=2e-5, atol=1e-12)      # csv carries 6 significant digits
        assert np.allclose(rows[:, 5], oracle.symmetry_error(ref["results"]), rtol=1e-3, atol=1e-12)
    assert (tmp_path / "OriginalMesh.obj").exists() and (tmp_path / "OriginalMesh.vtp").exists()
    assert open(tmp_path / "OriginalMesh.obj").read().startswith("v ")


@pytest.mark.gpu
def test_cli_adaptive_lines_match_reference_log(tmp_path):
    """stdout of the adaptive run on G1 against what the reference printed on the B200 (golden log)."""
    from test_golden_reference import META
    m = load_fixture("G1")
    write_dat(str(tmp_path / "G1.dat"), m)
    r = _run(["-f", "G1.dat", "--exportresults", "--exporttovtk"], tmp_path)
    assert r.returncode == 0, r.stderr
    mine = [ln for ln in r.stdout.splitlines() if ln.startswith(("Iteration", "Out of", "Integrating over", "Using adaptive", "Found"))]
    ref = [ln for ln in META["G1_ad"]["log"] if ln.startswith(("Iteration", "Out of", "Integrating over", "Using adaptive", "Found"))]
    assert "Using adaptive error control procedure" in r.stdout
    assert [x for x in mine if not x.startswith("Using")] == [x for x in ref if not x.startswith("Using")]
    first = open(tmp_path / "NotNeighbors.dat").readline()
    assert re.match(r"\(\d+, \d+\): \[[-0-9.e]+, [-0-9.e]+, [-0-9.e]+\]$", first.strip()), first
    vtp = open(tmp_path / "OriginalMesh.vtp").read()
    assert 'Name="NotNeighborsRefinements"' in vtp and 'Name="SimpleNeighborsRefinements"' in vtp


@pytest.mark.gpu
def test_cli_fixed_refinement_exports_refined_mesh(tmp_path):
    m = load_fixture("Case-7-2")
    write_dat(str(tmp_path / "c.dat"), m)
    r = _run(["--meshfile=c.dat", "--refine=2", "--exporttoobj"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Using fixed refinement level equal to 2" in r.stdout
    assert "Refined mesh contains 35 vertices and 32 cells. Number of tasks: simple neighbors - 32, attached neighbors - 0, non-neighbors - 0" in r.stdout
    obj = open(tmp_path / "RefinedMesh.obj").read().splitlines()
    assert sum(1 for ln in obj if ln.startswith("v ")) == 35 and sum(1 for ln in obj if ln.startswith("f ")) == 32

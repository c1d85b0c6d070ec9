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
    assert "--exportbinary" in r.stdout and "I2_GPUS" in r.stdout      # round-2 additions: binary export, multi-GPU environment
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


def read_binary_export(path):
    """<Name>.bin written by --exportbinary (Evaluator3D::outputResultsToFile, output_format_enum::binary)"""
    raw = open(path, "rb").read()
    assert raw[:4] == b"I2RB"
    version, = np.frombuffer(raw, dtype=np.int32, count=1, offset=4)
    n, = np.frombuffer(raw, dtype=np.int64, count=1, offset=8)
    has_err, _ = np.frombuffer(raw, dtype=np.int32, count=2, offset=16)
    assert version == 1
    off = 24
    tasks = np.frombuffer(raw, dtype=np.int32, count=3 * n, offset=off).reshape(n, 3); off += 12 * n
    results = np.frombuffer(raw, dtype=np.float64, count=3 * n, offset=off).reshape(n, 3); off += 24 * n
    errors = np.frombuffer(raw, dtype=np.float64, count=n, offset=off) if has_err else None
    assert len(raw) == off + (8 * n if has_err else 0)
    return tasks, results, errors


@pytest.mark.gpu
def test_cli_binary_export_and_defect_summary(tmp_path, oracle):
    """--exportbinary: full-precision records keyed (i, j) (the csv keeps 6 digits); -c prints one summary line per
    class (the reference computes the defects and prints nothing, D9); I2_SUMMARY_JSON writes the run summary."""
    import json
    m = load_fixture("G1")
    write_dat(str(tmp_path / "G1.dat"), m)
    env = dict(os.environ, I2_SUMMARY_JSON=str(tmp_path / "summary.json"))
    r = subprocess.run([CLI, "-f", "G1.dat", "-r", "0", "-c", "--exportbinary"], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    om = oracle.OracleMesh(m.vertices, m.cells)
    js = json.load(open(tmp_path / "summary.json"))
    assert js["gpus"] == 1 and js["level"] == 0 and [c["tasks"] for c in js["classes"]] == [898, 318, 9914]
    names = ("simple neighbors", "attached neighbors", "not neighbors")
    for cls, fn in enumerate(("SimpleNeighbors.bin", "AttachedNeighbors.bin", "NotNeighbors.bin")):
        tasks, results, errors = read_binary_export(tmp_path / fn)
        assert np.array_equal(tasks, om.tasks(cls))
        ref = om.run_class(cls, om.tasks(cls), 0)
        assert (np.abs(results - ref["results"]).sum(1) <= 1e-12 * np.abs(ref["results"]).sum(1)).all()
        d = oracle.symmetry_error(results)
        assert np.array_equal(errors, d)                 # the defect is a function of the stored results: bit-exact
        mline = re.search(rf"Symmetry check \(i,j\)/\(j,i\) for {names[cls]}: max delta = ([-0-9.e+]+), mean delta = ([-0-9.e+]+) \((\d+) ordered pairs\)", r.stdout)
        assert mline, r.stdout
        assert float(mline.group(1)) == pytest.approx(d.max(), rel=1e-5) and float(mline.group(2)) == pytest.approx(d.mean(), rel=1e-5)
        assert int(mline.group(3)) == tasks.shape[0]
        assert js["classes"][cls]["delta_max"] == pytest.approx(d.max(), rel=1e-4, abs=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("cls_file", ["SimpleNeighbors", "AttachedNeighbors", "NotNeighbors"])
def test_cli_csv_is_byte_identical_to_the_reference_file(tmp_path, cls_file):
    """The csv the REFERENCE's own CLI wrote for G1 `-r 0 --exporttocsv` on a B200 (rows sorted by (i, j): its row order is atomicAdd
    order; tools/gpu_golden_csv.sh produced tests/golden/G1_r0_csv.npz) against the drop-in's file, sorted the same way: equal bytes."""
    golden = os.path.join(ROOT, "tests", "golden", "G1_r0_csv.npz")
    if not os.path.exists(golden):
        pytest.skip("tests/golden/G1_r0_csv.npz not generated yet (tools/gpu_golden_csv.sh)")
    ref_text = bytes(np.load(golden)[cls_file]).decode()
    m = load_fixture("G1")
    write_dat(str(tmp_path / "G1.dat"), m)
    r = _run(["-f", "G1.dat", "-r", "0", "--exporttocsv"], tmp_path)
    assert r.returncode == 0, r.stderr

    def canon(text):
        lines = text.splitlines()
        return "\n".join([lines[0]] + sorted(lines[1:], key=lambda ln: tuple(int(x) for x in ln.split(";")[:2])))
    mine = canon(open(tmp_path / (cls_file + ".csv")).read())
    assert mine == canon(ref_text)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [["-r", "0"], ["-r", "1"], []])
def test_cli_two_gpus_writes_the_same_files(tmp_path, flags):
    """env I2_GPUS=2: Evaluator3D::runAllPairs shards the lists over two GPUs through i2_mgpu_* (sharded prepare, NCCL inside the
    library); csv, binary and vtp exports are byte-identical to the single-GPU run's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2; log in profiles/r02_mgpu_2gpu.txt)")
    m = load_fixture("s5m", 0.0005)
    outs = {}
    for gpus in (1, 2):
        d = tmp_path / f"g{gpus}"
        d.mkdir()
        write_dat(str(d / "m.dat"), m)
        env = dict(os.environ, I2_GPUS=str(gpus))
        r = subprocess.run([CLI, "-f", "m.dat", "-c", "--exporttocsv", "--exportbinary", "--exporttovtk"] + flags, cwd=d, capture_output=True, text=True,
                           timeout=900, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[gpus] = r.stdout
    assert "on 2 GPUs" in outs[2]
    for fn in ("SimpleNeighbors.csv", "AttachedNeighbors.csv", "NotNeighbors.csv", "SimpleNeighbors.bin", "AttachedNeighbors.bin", "NotNeighbors.bin",
               "OriginalMesh.vtp"):
        a, b = open(tmp_path / "g1" / fn, "rb").read(), open(tmp_path / "g2" / fn, "rb").read()
        assert a == b, fn
    pick = lambda out: [ln for ln in out.splitlines() if ln.startswith(("Out of", "Symmetry check"))]      # noqa: E731
    assert pick(outs[1]) == pick(outs[2])

"""CPU-side checks of bench.py: pair counts without a GPU, and the reference arm's fallback (the CPU oracle port on a bounded
sample when the reference's CUDA build cannot run) printing the contract's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from integrator2_b200.meshio import load_fixture

sys.path.insert(0, ROOT)


@pytest.mark.parametrize("name,scale,expect", [("G1", 1.0, (898, 318, 9914)), ("s5m", 0.0005, (19304, 5592, 3447736)),
                                                ("Vint16k", 1.0, (153530, 50790, 286403650)), ("s5m2", 0.0005, (75398, 23490, 61202182))])
def test_class_pair_counts_from_incidence(name, scale, expect):
    """SURVEY.md 8(d) task counts of the BASELINE.json configs, reproduced from vertex / edge incidence alone."""
    import bench
    assert bench.class_pair_counts(load_fixture(name, scale)) == expect


def test_class_pair_counts_match_the_oracle_classification(oracle):
    import bench
    for name in ("cubehole", "ellipsoid2000", "1x1x1_extrafine"):
        m = load_fixture(name)
        om = oracle.OracleMesh(m.vertices, m.cells)
        assert bench.class_pair_counts(m) == tuple(2 * int(x.shape[0]) for x in om.classify()), name


def test_reference_arm_prints_the_contract_line_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU fallback of the reference arm: only meaningful without a GPU")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--mesh", "G1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["config"]["workload"].startswith("G1.dat scale 1.0 level 0: 898 vertex-adjacent + 318 edge-adjacent + 9914 regular")
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert np.isfinite(line["value"]) and line["value"] > 0


def test_reference_command_never_passes_a_negative_level():
    """Automatic error control is the ABSENCE of -r in the reference CLI (tests/integrator3D/main.cu:114-123)."""
    import bench
    assert bench.reference_command("bin", "m.dat", 1.0, 0) == ["bin", "-f", "m.dat", "-r", "0"]
    assert bench.reference_command("bin", "m.dat", 0.0005, 2) == ["bin", "-f", "m.dat", "-s", "0.0005", "-r", "2"]
    adaptive = bench.reference_command("bin", "m.dat", 0.0005, -1)
    assert "-r" not in adaptive and adaptive == ["bin", "-f", "m.dat", "-s", "0.0005"]

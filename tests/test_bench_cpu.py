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


def test_bench_reads_the_kernel_counters_from_the_committed_ncu_summary():
    """bench.py's `roofline.counters` / `executed` come from profiles/r02_ncu_k_regular_grouped*.json (the newest one), never
    from literals: the file must exist, name the kernel and carry every key bench.py reads."""
    import bench
    prof = bench.load_kernel_profile()
    assert prof, "no committed ncu summary of the regular-pair kernel"
    assert prof["source"].startswith("profiles/r02_ncu_k_regular_grouped") and os.path.exists(os.path.join(ROOT, prof["source"]))
    assert "k_regular_grouped" in prof["kernel"]
    for key in ("fp64_inst_per_pair", "other_warp_inst_per_pair", "fp64_pipe_active_frac", "issue_slots_active_frac",
                "achieved_occupancy_frac", "registers", "stall_wait_frac", "stall_math_throttle_frac", "dram_bytes_read",
                "dram_bytes_write", "pairs_per_launch"):
        assert key in prof, key
    assert 1000 < prof["fp64_inst_per_pair"] < 1300 and 0.5 < prof["fp64_pipe_active_frac"] < 1.0
    # the DRAM traffic of the capture is what bench.py reports as roofline.traffic: close to the algorithmic 68 B per pair
    per_pair = (prof["dram_bytes_read"] + prof["dram_bytes_write"]) / prof["pairs_per_launch"]
    assert 56.0 < per_pair < 80.0


def test_ncu_summary_tool_on_a_synthetic_raw_page(tmp_path):
    """tools/ncu_summary.py: per-pair instruction counts = per-cycle rates x elapsed cycles / pairs, units honoured."""
    hdr = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
           "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    units = ["", "", "", "", "us", "register/thread", "cycle", "inst", "inst/cycle", "inst/cycle", "inst/cycle", "%", "Mbyte", "Kbyte",
             "", ""]
    row = ["0", "void i2::k_demo<4>(int)", "(148, 1, 1)", "(128, 1, 1)", "2,000", "96", "1000", "5,000", "600", "300", "100", "62.5",
           "3", "500", "3.0", "1.0"]
    other = ["1", "k_other(int)", "(1, 1, 1)", "(32, 1, 1)", "1", "16", "10", "1", "0", "0", "0", "0", "0", "0", "1.0", "0.0"]
    raw = tmp_path / "raw.csv"
    import csv
    with open(raw, "w", newline="") as f:
        csv.writer(f, quoting=csv.QUOTE_ALL).writerows([hdr, units, row, other])
    out = tmp_path / "summary"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), str(raw), "--kernel", "k_demo", "--pairs", "1000",
                        "--out", str(out), "--note", "synthetic"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = json.load(open(str(out) + ".json"))
    assert d["duration_ms_under_ncu"] == pytest.approx(2.0)            # 2 000 us
    assert d["fp64_inst_per_pair"] == pytest.approx(1000.0)             # (600 + 300 + 100) per cycle x 1000 cycles / 1000 pairs
    assert d["all_inst_per_pair"] == pytest.approx(160.0)               # 5 000 warp instructions x 32 lanes / 1000 pairs
    assert d["fp64_pipe_active_frac"] == pytest.approx(0.625)
    assert d["dram_bytes_per_pair"] == pytest.approx(3500.0)            # (3 MB + 500 kB) / 1000
    assert d["stall_wait_frac"] == pytest.approx(0.75) and d["stall_math_throttle_frac"] == pytest.approx(0.25)
    assert "k_demo" in d["kernel"] and os.path.exists(str(out) + ".txt")
    # a kernel that is not in the capture is an error, not an empty summary
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), str(raw), "--kernel", "k_missing", "--pairs", "1",
                        "--out", str(out)], capture_output=True, text=True)
    assert r.returncode != 0

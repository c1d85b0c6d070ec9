#!/bin/bash
# list-free adaptive kernel: tests + timing on s5m / Vint16k rows
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matrix_free.py -m gpu -q -x -k adaptive --timeout 500 > gpurun_out/pytest_adfree.log 2>&1; tail -30 gpurun_out/pytest_adfree.log
timeout 300 python - <<'PY' 2>&1 | tail -12
import torch, time, json
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
ctx = abi.Context(0)
for name, scale, rows in (("s5m", 0.0005, None), ("s5m2", 0.0005, None), ("Vint16k", 1.0, 4096)):
    m = load_fixture(name, scale)
    ctx.set_mesh(m.vertices, m.cells)
    hi = rows or m.n_cells
    ctx.apply_regular_adaptive(0, hi)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = ctx.apply_regular_adaptive(0, hi)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.apply_regular(0, hi); torch.cuda.synchronize(); d0 = time.perf_counter() - t0
    print(json.dumps({"mesh": name, "rows": hi, "adaptive_ms": dt * 1e3, "level0_ms": d0 * 1e3, "pairs": a["stats"]["integrated"][0],
                      "pairs_per_s": a["stats"]["integrated"][0] / dt, "stats": a["stats"], "refinement_hist": torch.bincount(a["refinements"].int()).tolist()}))
PY

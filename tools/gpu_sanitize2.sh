#!/bin/bash
# round 2: compute-sanitizer over the NEW kernels — classification by vertex incidence, sharded list fill, halves-aware warp groups,
# per-task vote masks (adaptive round 2), coalesced remote-style stores, row-major adjacent lists + row scatter, operator apply
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
for name, scale in (("G1", 1.0), ("cubehole", 1.0)):
    m = load_fixture(name, scale)
    whole = abi.Context(0)
    counts = whole.host_prepare(m.vertices, m.cells)
    for level in (0, 1, -1):
        whole.host_run(level, None, None)
    shards = []
    for r in range(3):
        c = abi.Context(0)
        c.host_set_shard(r, 3)
        c.host_prepare(m.vertices, m.cells)
        shards.append(c)
    for level in (0, -1):
        for c in shards:
            c.host_run_rounds(level)
        if level < 0:
            L = np.max([c.host_last_rounds() for c in shards], axis=0).tolist()
            ref = np.max([c.host_refinements() for c in shards], axis=0)
            for c in shards:
                c.host_last_rounds(L); c.host_refinements(ref)
        for c in shards:
            c.host_run_finalize(level, check=True)
            for k in range(3):
                c.host_fetch(k, errors=True)
    ctx = abi.Context(0)
    ctx.set_mesh(m.vertices, m.cells)
    lists = ctx.classify()
    ctx.apply_prepare(0, m.n_cells)
    for level in (0, -1):
        a = ctx.apply(level)
        assert torch.isfinite(a["out"]).all()
    ctx.apply_prepare(32, min(96, m.n_cells))
    ctx.apply(-1)
    mg = abi.MultiGpu(local_gpus=1)
    mg.prepare(m.vertices, m.cells, -1)
    mg.run(-1, check=True, want_stats=True)
    mg.checksums()
    mg.apply_prepare(m.vertices, m.cells, -1)
    mg.apply(-1)
    mg.close()
    import os
    os.environ["I2_VEC_STORES"] = "1"
print("sanitizer workload done")
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer2_$tool.log python /tmp/san2.py > gpurun_out/sanitizer2_$tool.out 2>&1
  tail -1 gpurun_out/sanitizer2_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer2_$tool.log | head -3
done
echo "== coalesced stores forced on (I2_VEC_STORES=1), memcheck"
I2_VEC_STORES=1 timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer2_vec.log python /tmp/san2.py > gpurun_out/sanitizer2_vec.out 2>&1
tail -1 gpurun_out/sanitizer2_vec.out; grep -E "ERROR SUMMARY" gpurun_out/sanitizer2_vec.log | head -2

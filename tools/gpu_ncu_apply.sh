#!/bin/bash
# ncu --set full of the list-free level-0 row-sum kernel at 5 CTAs per SM (k_apply_regular<5,16>) on s5m2 refined once (31 320 triangles)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cat > /tmp/prof_apply.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture, subdivide
m = subdivide(load_fixture("s5m2", 0.0005), 1)
c = abi.Context(0)
c.set_mesh(m.vertices, m.cells)
for _ in range(3):
    out = c.apply_regular(0, len(m.cells))
c.synchronize()
print("triangles", len(m.cells), "checksum", float(out.abs().sum()))
PY
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:^k_apply_regular$' -s 2 -c 1 -f -o gpurun_out/n_prof_apply python /tmp/prof_apply.py > gpurun_out/n_ncu_apply.log 2>&1
tail -3 gpurun_out/n_ncu_apply.log
ls -la gpurun_out/n_prof_apply.ncu-rep

#!/bin/bash
# one ncu --set full capture of the list-free adaptive kernel (s5m2, 2048 rows x 7830 columns)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/adfree_run.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
ctx = abi.Context(0)
m = load_fixture("s5m2", 0.0005)
ctx.set_mesh(m.vertices, m.cells)
for _ in range(2):
    a = ctx.apply_regular_adaptive(0, 2048)
torch.cuda.synchronize()
print(a["stats"])
PY
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_apply_regular_adaptive -s 1 -c 1 -o gpurun_out/prof_apply_adaptive python /tmp/adfree_run.py > gpurun_out/ncu_adfree.log 2>&1
tail -2 gpurun_out/ncu_adfree.log

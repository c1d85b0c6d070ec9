#!/bin/bash
# ncu --set full of the two other FP64-heavy kernels on the final binary: the list-free adaptive kernel (largest-mesh path) and the
# LEVEL1 specialisation of the regular kernel (round 1 of the work queue), both on s5m2 (61.2 M regular pairs)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cat > /tmp/prof_extra.py <<'PY'
import sys
sys.path.insert(0, ".")
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
m = load_fixture("s5m2", 0.0005)
c = abi.Context(0)
if sys.argv[1] == "apply":
    c.set_mesh(m.vertices, m.cells)
    for _ in range(2):
        a = c.apply_regular_adaptive(0, m.n_cells)
    print(a["stats"])
else:
    c.host_prepare(m.vertices, m.cells)
    for _ in range(2):
        c.host_run_rounds(-1); c.host_run_finalize(-1)
c.synchronize()
PY
true
true
I2_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_regular_grouped -s 9 -c 1 -f -o gpurun_out/x_prof_level1 python /tmp/prof_extra.py lists > gpurun_out/x_ncu_level1.log 2>&1
tail -2 gpurun_out/x_ncu_level1.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cat > /tmp/ad_child.py <<'PY'
import sys
sys.path.insert(0, ".")
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
name, scale = sys.argv[1], float(sys.argv[2])
m = load_fixture(name, scale)
c = abi.Context(0)
c.host_prepare(m.vertices, m.cells)
for _ in range(3):
    c.host_run_rounds(-1); c.host_run_finalize(-1)
c.synchronize()
PY
for mesh in s5m s5m2; do
I2_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j_launches_${mesh}_adaptive.csv python /tmp/ad_child.py $mesh 0.0005 > gpurun_out/j_ncu_$mesh.log 2>&1
tail -1 gpurun_out/j_ncu_$mesh.log
done

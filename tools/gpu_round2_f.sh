#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/f_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/f_all_tests.log
tail -8 gpurun_out/f_all_tests.log
timeout 900 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
tail -c 800 gpurun_out/f_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
cat > /tmp/prof_child.py <<'PY'
import sys
sys.path.insert(0, ".")
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
m = load_fixture("Vint16k")
c = abi.Context(0)
c.host_prepare(m.vertices, m.cells)
for _ in range(3):
    c.host_run_rounds(0)
c.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_regular_grouped -s 2 -c 1 -f -o gpurun_out/f_prof_grouped python /tmp/prof_child.py > gpurun_out/f_ncu_full.log 2>&1
tail -2 gpurun_out/f_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-largest > gpurun_out/f_ncu_launches.log 2>&1
cd /tmp && python - <<'PY'
import sys
sys.path.insert(0, "/root/repo")
from integrator2_b200.meshio import load_fixture, write_dat
write_dat("/tmp/Vint16k.dat", load_fixture("Vint16k"))
PY
"$GRAFT_REPO_ROOT/integrator2_b200/host/integrator2test3D" -f /tmp/Vint16k.dat -r 0 -c > "$GRAFT_REPO_ROOT/gpurun_out/f_cli_gpus1.txt" 2>&1
grep -E "Time for|Symmetry" "$GRAFT_REPO_ROOT/gpurun_out/f_cli_gpus1.txt"

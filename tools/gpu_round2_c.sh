#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tools/gpu_ab.py BASE=1 I2_MINBLOCKS=5 I2_VARIANT=59 I2_MINBLOCKS=5,I2_VARIANT=59 I2_MINBLOCKS=3 I2_VEC_STORES=1 BASE=1 > gpurun_out/c_ab.log 2>&1
cat gpurun_out/c_ab.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/c_all_tests.log
tail -12 gpurun_out/c_all_tests.log
timeout 900 python bench.py > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err
tail -c 1500 gpurun_out/c_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err
tail -c 600 gpurun_out/c_bench_ref.json
# ncu: full capture of the regular-pair kernel + launch list of a short bench run
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_regular_grouped -s 2 -c 1 -f -o gpurun_out/c_prof_grouped python /tmp/prof_child.py > gpurun_out/c_ncu_full.log 2>&1
tail -3 gpurun_out/c_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-largest > gpurun_out/c_ncu_launches.log 2>&1
tail -2 gpurun_out/c_ncu_launches.log

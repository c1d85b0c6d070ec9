#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
bash tools/gpu_golden_csv.sh > gpurun_out/b_golden_csv.log 2>&1
python tools/gpu_ab.py I2_VEC_STORES=1 I2_VEC_STORES=0 I2_VEC_STORES=1 I2_VEC_STORES=0 > gpurun_out/b_ab_stores.log 2>&1
cat gpurun_out/b_ab_stores.log
timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_apply.py::test_largest_meshes_against_the_oracle_on_sampled_rows > gpurun_out/b_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/b_all_tests.log
tail -25 gpurun_out/b_all_tests.log
timeout 1200 python -m pytest tests/test_gpu_apply.py -m gpu -q -k largest > gpurun_out/b_largest.log 2>&1
tail -25 gpurun_out/b_largest.log

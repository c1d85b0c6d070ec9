#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tools/gpu_ab.py I2_FIX13=0 I2_FIX13=1 I2_FIX13=1,I2_VARIANT=59 I2_FIX13=0 I2_FIX13=1 > gpurun_out/e_ab_fix13.log 2>&1
cat gpurun_out/e_ab_fix13.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_mgpu.py tests/test_gpu_quadrature_rules.py tests/test_gpu_apply.py -m gpu -q -x > gpurun_out/e_tests.log 2>&1
echo "exit $?" >> gpurun_out/e_tests.log
tail -8 gpurun_out/e_tests.log

#!/bin/bash
# after switching the list-free kernels to 5 / 4 CTAs per SM: the tests that reach them, smoke, and the default bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_apply.py tests/test_gpu_matrix_free.py tests/test_gpu_mgpu.py -m gpu -q > gpurun_out/n_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/n_tests.log
tail -4 gpurun_out/n_tests.log
timeout 60 python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/n_smoke.log 2>&1
tail -1 gpurun_out/n_smoke.log
timeout 170 python bench.py > gpurun_out/n_bench_n1.json 2> gpurun_out/n_bench_n1.err
echo "bench exit $?"
tail -c 1200 gpurun_out/n_bench_n1.json

#!/bin/bash
# round 2, first GPU pass: new tests first (fast feedback), then the whole GPU suite, then the bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt
timeout 1500 python -m pytest tests/test_gpu_mgpu.py tests/test_gpu_apply.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/a_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/a_new_tests.log
tail -30 gpurun_out/a_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/a_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/a_all_tests.log
tail -15 gpurun_out/a_all_tests.log
timeout 900 python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
tail -c 3000 gpurun_out/a_bench_n1.json
tail -5 gpurun_out/a_bench_n1.err

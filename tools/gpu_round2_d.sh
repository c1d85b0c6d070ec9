#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/d_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/d_all_tests.log
tail -12 gpurun_out/d_all_tests.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/d_smoke.log 2>&1
tail -2 gpurun_out/d_smoke.log

#!/bin/bash
# 2-GPU pass (gpurun --gpus 2): the multi-GPU tests the driver's 1-GPU lease skips, kept as a log under profiles/
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mgpu_2gpu.txt
timeout 1800 python -m pytest tests/test_gpu_mgpu.py tests/test_cli.py tests/test_gpu_peer_export.py -m gpu -v >> gpurun_out/r02_mgpu_2gpu.txt 2>&1
echo "exit $?" >> gpurun_out/r02_mgpu_2gpu.txt
grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed" gpurun_out/r02_mgpu_2gpu.txt | tail -40
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR --nproc-per-node 2 bench.py --gpus 2 --no-cpu > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err
tail -c 2500 gpurun_out/d_bench_n2.json; tail -3 gpurun_out/d_bench_n2.err
# configs[3] as BASELINE.json states it (s5m2.dat, automatic error control, lists), N = 1 and 2
timeout 600 python bench.py --mesh s5m2 --scale 0.0005 --level -1 --no-cpu --no-largest > gpurun_out/d_bench_s5m2_ad_n1.json 2> gpurun_out/d_bench_s5m2_ad_n1.err
timeout 600 $TR --nproc-per-node 2 bench.py --gpus 2 --mesh s5m2 --scale 0.0005 --level -1 --no-cpu --no-largest > gpurun_out/d_bench_s5m2_ad_n2.json 2> gpurun_out/d_bench_s5m2_ad_n2.err
for f in gpurun_out/d_bench_s5m2_ad_n1.json gpurun_out/d_bench_s5m2_ad_n2.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -3 gpurun_out/d_bench_s5m2_ad_n2.err

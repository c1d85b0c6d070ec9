#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29513"
timeout 400 $TR --nproc-per-node 8 bench.py --gpus 8 --no-cpu > gpurun_out/e_bench_n8.json 2> gpurun_out/e_bench_n8.err
timeout 400 python bench.py --no-cpu > gpurun_out/e_bench_n1_samebox.json 2> gpurun_out/e_bench_n1_samebox.err
for f in gpurun_out/e_bench_n8.json gpurun_out/e_bench_n1_samebox.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    lm = d.get("largest_mesh") or {}
    print(sys.argv[1], d["n_gpus"], "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "largest ms", lm.get("ms_per_step"), "sphere ms", (lm.get("sphere") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done

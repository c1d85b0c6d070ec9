#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for p in 0 1; do for cfg in "s5m 0.0005 -1" "s5m2 0.0005 -1" "Vint16k 1.0 0"; do
set -- $cfg
I2_SIDE_PRIORITY=$p timeout 300 python bench.py --mesh $1 --scale $2 --level $3 --no-cpu --no-largest --steps 10 > gpurun_out/i_bench_$1_prio$p.json 2> gpurun_out/i_bench_$1_prio$p.err
python - gpurun_out/i_bench_$1_prio$p.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "kernel", (d.get("roofline") or {}).get("kernel_ms"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done; done

#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): N = 8 and N = 1 back to back on the same box, CLI with I2_GPUS=8
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/e_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR --nproc-per-node 8 bench.py --gpus 8 --no-cpu > gpurun_out/e_bench_n8.json 2> gpurun_out/e_bench_n8.err
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --mesh s5m2 --scale 0.0005 --level -1 --no-cpu --no-largest > gpurun_out/e_bench_s5m2_ad_n8.json 2> gpurun_out/e_bench_s5m2_ad_n8.err
# the drop-in CLI on 8 GPUs vs 1 (its own wall-clock line)
python - <<'PY'
import sys
sys.path.insert(0, ".")
from integrator2_b200.meshio import load_fixture, write_dat
write_dat("/tmp/Vint16k.dat", load_fixture("Vint16k"))
PY
cd /tmp
for g in 8; do I2_GPUS=$g timeout 600 "$GRAFT_REPO_ROOT/integrator2_b200/host/integrator2test3D" -f /tmp/Vint16k.dat -r 0 -c > "$GRAFT_REPO_ROOT/gpurun_out/e_cli_gpus$g.txt" 2>&1; done
cd "$GRAFT_REPO_ROOT"
grep -E "Time for|Symmetry" gpurun_out/e_cli_gpus8.txt
for f in gpurun_out/e_bench_n8.json gpurun_out/e_bench_s5m2_ad_n8.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    lm = d.get("largest_mesh") or {}
    print("per rank ms", [round(x, 3) for x in (d.get("per_rank_ms_per_step") or [])])
    print(sys.argv[1], d["n_gpus"], "ms", round(d["ms_per_step"], 3), "e2e ms", (d.get("e2e") or {}).get("ms_per_step"), "largest ms", lm.get("ms_per_step"),
          "gather", (d.get("with_gather_to_rank0") or {}).get("ms_per_step"), "peer", (d.get("with_peer_store_to_rank0") or {}).get("ms_per_step"),
          "ingest GB/s", (d.get("nvlink_ingest") or {}).get("gb_per_s"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done

#!/bin/bash
# last check of the driver-facing commands on a 2-GPU box: one JSON line on stdout at N = 1 and N = 2, reference arm
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29514"
timeout 400 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/y_bench_n2.json 2> gpurun_out/y_bench_n2.err
timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/y_bench_n1.json 2> gpurun_out/y_bench_n1.err
timeout 300 $TR --nproc-per-node 2 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/y_bench_ref_n2.json 2> gpurun_out/y_bench_ref_n2.err
for f in gpurun_out/y_bench_n2.json gpurun_out/y_bench_n1.json gpurun_out/y_bench_ref_n2.json; do
  echo "$f: $(wc -l < $f) line(s)"; python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print("  ", d.get("impl", "ours"), "N", d["n_gpus"], "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), "G pairs/s",
      "export err", d.get("export_variants_error"), "largest", (d.get("largest_mesh") or {}).get("ms_per_step"))
PY
done

"""Exploration on the GPU box: kernel timings per variant and parity statistics (prints JSON lines)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture

what = sys.argv[1] if len(sys.argv) > 1 else "time"
mesh_name = sys.argv[2] if len(sys.argv) > 2 else "Vint16k"
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0

ctx = abi.Context(0)
m = load_fixture(mesh_name, scale)
ctx.set_mesh(m.vertices, m.cells)
lists = ctx.classify()
tasks = [ctx.tasks_from_pairs(p) for p in lists]
counts = [int(t.shape[0]) for t in tasks]

if what == "time":
    ctx.set_profiling(True)
    out = {"mesh": mesh_name, "counts": counts, "minblocks": os.environ.get("I2_MINBLOCKS", "4")}
    for mode in (1, 3, 2, 0):
        ctx.set_math_mode(mode)
        for cls in (2, 1, 0):
            if mode != 1 and cls != 2:
                continue
            n = counts[cls]
            buf = (torch.empty((n, 4), dtype=torch.float64, device="cuda"), torch.empty((n, 3), dtype=torch.float64, device="cuda"))
            ts = []
            for rep in range(4):
                ctx.integrate_class(cls, tasks[cls], 0, want_stats=False, out=buf)
                ts.append(ctx.profile_last())
            ti, tf = min(t[0] for t in ts[1:]), min(t[1] for t in ts[1:])
            out[f"cls{cls}_mode{mode}"] = {"integrate_ms": ti, "finalize_ms": tf, "pairs_per_s": n / (ti * 1e-3)}
            del buf
    out["peak_dfma_tflops"], out["peak_mufu_gops"] = ctx.peak_rates()
    print(json.dumps(out))
elif what == "adaptive":
    for cls in (0, 1, 2):
        ctx.integrate_class(cls, tasks[cls], -1)     # warm-up (allocations, kernel images)
        torch.cuda.synchronize()
        t0 = time.time()
        r = ctx.integrate_class(cls, tasks[cls], -1)
        torch.cuda.synchronize()
        print(json.dumps({"mesh": mesh_name, "cls": cls, "adaptive_s": time.time() - t0, "stats": r["stats"],
                          "refinement_hist": np.bincount(r["refinements"].cpu().numpy()).tolist()}))

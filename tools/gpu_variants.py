"""Times the regular-pair kernel for the tuning knobs given in the environment (I2_MINBLOCKS, I2_VARIANT)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture

if os.environ.get("I2_LIB_PATH"):       # an experimental build of the library (tools/gpu_unroll.sh)
    abi.LIB_PATH = os.path.abspath(os.environ["I2_LIB_PATH"])

ctx = abi.Context(0)
m = load_fixture("Vint16k")
ctx.set_mesh(m.vertices, m.cells)
lists = ctx.classify()
tasks = ctx.tasks_from_pairs(lists[2])
n = int(tasks.shape[0])
buf = (torch.empty((n, 4), dtype=torch.float64, device="cuda"), torch.empty((n, 3), dtype=torch.float64, device="cuda"))
ctx.set_profiling(True)
out = {"lib": os.path.basename(abi.LIB_PATH), "minblocks": os.environ.get("I2_MINBLOCKS", "4"), "variant": os.environ.get("I2_VARIANT", "0")}
ts = []
for rep in range(5):
    ctx.integrate_class(2, tasks, 0, want_stats=False, out=buf)
    ts.append(ctx.profile_last()[0])
out["level0_ms"] = min(ts[1:])
out["pairs_per_s"] = n / (out["level0_ms"] * 1e-3)
sub = tasks[: n // 8].contiguous()
ts = []
for rep in range(3):
    ctx.integrate_class(2, sub, 1, want_stats=False, out=(buf[0][: n // 8], buf[1][: n // 8]))
    ts.append(ctx.profile_last()[0])
out["level1_ms_per_8th"] = min(ts[1:])
out["checksum"] = float(buf[1][: n // 8].abs().sum())
out["peak_dfma_tflops"], out["peak_mufu_gops"] = ctx.peak_rates()
out["peak_dfma3_tflops"] = ctx.peak_dfma_three_operand()
print(json.dumps(out))

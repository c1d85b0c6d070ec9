"""A/B timing of the list-free level-0 row-sum kernel (k_apply_regular) on the configs[4] sphere (108 544 triangles):
one child process per setting of the env knobs on the command line, e.g.  python tools/gpu_ab_apply.py I2_APPLY_MINB=4 I2_APPLY_MINB=5"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, sys.argv[1])
import torch
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture, subdivide
m = subdivide(load_fixture("G1", 1.0), 5)
n = len(m.cells)
c = abi.Context(0)
c.set_mesh(m.vertices, m.cells)
out = torch.empty((n, 3), dtype=torch.float64, device="cuda:0")
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    c.apply_regular(0, n, None, out)
    c.synchronize(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
digest = int(out.view(torch.int64).sum().item()) & 0xffffffffffffffff
print("k_apply_regular ms: min %.2f median %.2f  row-sum bits digest %016x" % (min(ts[1:]), sorted(ts[1:])[len(ts[1:]) // 2], digest), os.environ.get("AB_LABEL"))
# the list-free Runge loop (k_apply_regular_adaptive) on s5m2 refined once: 31 320 triangles
m = subdivide(load_fixture("s5m2", 0.0005), 1)
n = len(m.cells)
c.set_mesh(m.vertices, m.cells)
ts = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    r = c.apply_regular_adaptive(0, n, None, want_stats=False)
    c.synchronize(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
digest = int(r["out"].view(torch.int64).sum().item()) & 0xffffffffffffffff
print("k_apply_regular_adaptive ms: min %.2f median %.2f  row-sum bits digest %016x" % (min(ts[1:]), sorted(ts[1:])[len(ts[1:]) // 2], digest), os.environ.get("AB_LABEL"))
'''
for setting in sys.argv[1:] or ["BASE=1"]:
    env = dict(os.environ, AB_LABEL=setting)
    for kv in setting.split(","):
        k, v = kv.split("=")
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-2000:], flush=True)

"""A/B timing of the regular-pair kernel on Vint16k level 0 (device events around i2_host_run_rounds): run once per setting of
the env knobs given on the command line, e.g.  python tools/gpu_ab.py I2_VEC_STORES=0 I2_VEC_STORES=1"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, sys.argv[1])
import torch
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
m = load_fixture("Vint16k")
c = abi.Context(0)
c.host_prepare(m.vertices, m.cells)
c.set_profiling(True)
ts = []
for _ in range(8):
    c.host_run_rounds(0)
    ts.append(c.profile_last()[0])
t_ptr, r_ptr = c.host_device_views()
n = c.host_shard()[1][2]
res = torch.as_tensor(abi._RawCudaBuffer(r_ptr[2], (n * 3,), "<i8"), device="cuda:0")
digest = int(res.sum().item()) & 0xffffffffffffffff          # wrap-around sum of the bit patterns: equal bits <=> equal digest (w.h.p.)
print("regular kernel ms: min %.3f median %.3f  result-bits digest %016x" % (min(ts[2:]), sorted(ts[2:])[len(ts[2:]) // 2], digest), os.environ.get("AB_LABEL"))
'''
for setting in sys.argv[1:] or ["BASE=1"]:
    env = dict(os.environ, AB_LABEL=setting)
    for kv in setting.split(","):
        k, v = kv.split("=")
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-2000:], flush=True)

MIX = r'''
import sys
sys.path.insert(0, sys.argv[1])
from integrator2_b200 import abi
c = abi.Context(0)
print("DFMA TFLOP/s with 0/1/2/3 independent LOP3 per DFMA:", [round(c.peak_dfma_with_integer(n), 2) for n in range(4)])
'''
print(subprocess.run([sys.executable, "-c", MIX, ROOT], capture_output=True, text=True).stdout.strip(), flush=True)

"""Pack the binary dumps of the REFERENCE's own CUDA build (written on the B200 box by tools/gpu_golden.sh with
oracle/_ref/ref_dump, i.e. the unmodified reference sources) into the small fixture tests/golden/reference_b200.npz.

Run here after `gpurun -- bash tools/gpu_golden.sh`:   python tools/pack_golden.py
Per dump <name> (e.g. G1_r0, s5m_ad, Case-7-2_r0) and class c in {simple, attached, not} the fixture holds
  <name>.<c>.tasks   int32 [n,3]   (i, j, k) as the reference ordered them (atomicAdd order) — compare keyed on (i,j)
  <name>.<c>.J       float64 [n,3] final J(K_i,K_j)  (Evaluator3D::d_*Results)
  <name>.<c>.I       float64 [n,4] (Psi, Theta)      (only for the small meshes)
  <name>.refinements uint8 [3,nc]  getRefinementsRequired (adaptive runs)
  <name>.log         the reference's stdout (iteration / convergence / timing lines)
Large classes are subsampled (every k-th record) to keep the file small; the subsampling is recorded in `meta`.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import read_class_dump, read_mesh_dump  # noqa: E402

SRC = os.path.join(ROOT, "gpurun_out", "golden")
OUT = os.path.join(ROOT, "tests", "golden", "reference_b200.npz")

names = sorted({f.split(".")[0] for f in os.listdir(SRC) if f.endswith(".mesh.bin")})
arrays, meta = {}, {}
for name in names:
    small_mesh = name.startswith(("Case", "G1", "Test", "genCase"))
    md = read_mesh_dump(os.path.join(SRC, name + ".mesh.bin"))
    if md["refinements"] is not None:
        arrays[name + ".refinements"] = np.stack(md["refinements"])
    meta[name] = {"nc": int(md["cells"].shape[0]), "adaptive": int(md["adaptive"]), "extra_stride": {}}
    for cname in ("simple", "attached", "not"):
        d = read_class_dump(os.path.join(SRC, f"{name}.{cname}.bin"))
        n = d["tasks"].shape[0]
        stride = 1
        if not small_mesh:
            stride = max(1, n // (6000 if cname == "not" else 4000))
        sl = slice(0, n, stride)
        arrays[f"{name}.{cname}.tasks"] = d["tasks"][sl]
        arrays[f"{name}.{cname}.J"] = d["results"][sl]
        if small_mesh and (not name.startswith("G1_") or name in ("G1_r0", "G1_ad")):
            arrays[f"{name}.{cname}.I"] = d["integrals"][sl]
        meta[name]["extra_stride"][cname] = stride
    log = os.path.join(SRC, name + ".log")
    if os.path.exists(log):
        keep = [ln for ln in open(log, errors="ignore").read().splitlines()
                if ln.startswith(("Iteration", "Out of", "Time for", "Found", "Loaded", "Integrating", "Orientation", "Refined mesh"))]
        meta[name]["log"] = keep
for cli in ("ref_cli_Vint16k_r0", "ref_cli_s5m2_ad"):
    p = os.path.join(SRC, cli + ".log")
    if os.path.exists(p):
        meta[cli] = {"log": [ln for ln in open(p, errors="ignore").read().splitlines()
                             if ln.startswith(("Iteration", "Out of", "Time for", "Found", "Loaded", "Integrating"))]}
arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
np.savez_compressed(OUT, **arrays)
print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(names), "dumps")

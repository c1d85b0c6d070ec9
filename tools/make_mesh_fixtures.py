"""Generate tests/golden/meshes.npz from the reference's example meshes.

Run in the build container (needs /root/reference):  python tools/make_mesh_fixtures.py
The fixture holds the parsed (unscaled) FP64 vertex coordinates and 0-based triangles of every
example mesh, i.e. input DATA only — no reference source code.  The GPU box has no /root/reference,
so tests and bench.py read this file.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from integrator2_b200.meshio import load_dat  # noqa: E402

SRC = "/root/reference/examples"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes.npz")

arrays = {}
for fn in sorted(os.listdir(SRC)):
    if fn in ("G1-contacti.dat", "G1-sosedi.dat"):  # legacy neighbour tables, not meshes
        continue
    try:
        m = load_dat(os.path.join(SRC, fn))
    except Exception as e:  # noqa: BLE001
        print("skip", fn, e)
        continue
    if m.n_cells == 0:  # G2-*.dat hold only line elements
        continue
    name = fn[:-4] if fn.endswith(".dat") else fn
    if name in arrays:
        continue
    if name + ".v" in arrays:
        continue
    arrays[name + ".v"] = m.vertices
    arrays[name + ".c"] = m.cells
    print(f"{name}: {m.vertices.shape[0]} vertices, {m.n_cells} triangles")
np.savez_compressed(OUT, **arrays)
print("wrote", OUT, os.path.getsize(OUT), "bytes")

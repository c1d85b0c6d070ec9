"""Mesh input in the reference's `.dat` format and the compact fixture format used by the tests.

Reference loader: /root/reference/src/Mesh3d.cu:159-217 — line 1 `nVertices nEntities`, then
`idx x y z` per vertex, then `idx type v1 v2 v3` with type 203 = triangle (1-based vertex ids);
any other entity is `idx type a b` and is skipped.  Coordinates are multiplied by `scale`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class TriMesh:
    vertices: np.ndarray  # float64 [nv, 3], already scaled
    cells: np.ndarray     # int32 [nc, 3], 0-based

    @property
    def n_cells(self) -> int:
        return int(self.cells.shape[0])


def parse_dat(text: str, scale: float = 1.0) -> TriMesh:
    tok = text.split()
    nv, ne = int(tok[0]), int(tok[1])
    pos = 2
    verts = np.empty((nv, 3), dtype=np.float64)
    for v in range(nv):
        verts[v, 0] = float(tok[pos + 1])
        verts[v, 1] = float(tok[pos + 2])
        verts[v, 2] = float(tok[pos + 3])
        pos += 4
    cells = []
    while pos + 1 < len(tok):
        etype = int(tok[pos + 1])
        if etype == 203:
            cells.append((int(tok[pos + 2]) - 1, int(tok[pos + 3]) - 1, int(tok[pos + 4]) - 1))
            pos += 5
        else:
            pos += 4
    # `scale * vertex` in the reference is a plain FP64 product per component (Mesh3d.cu:178)
    return TriMesh(verts * float(scale), np.asarray(cells, dtype=np.int32).reshape(-1, 3))


def load_dat(path: str, scale: float = 1.0) -> TriMesh:
    with open(path, "r") as f:
        return parse_dat(f.read(), scale)


def write_dat(path: str, mesh: TriMesh) -> None:
    """Write a mesh back in the reference's input format (round-trip exact: repr of doubles)."""
    with open(path, "w") as f:
        f.write(f"{mesh.vertices.shape[0]} {mesh.cells.shape[0]}\n")
        for k, v in enumerate(mesh.vertices):
            f.write(f"{k + 1} {float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")
        for k, c in enumerate(mesh.cells):
            f.write(f"{k + 1} 203 {int(c[0]) + 1} {int(c[1]) + 1} {int(c[2]) + 1}\n")


_FIXTURES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes.npz")
_fixture_cache = None


def fixture_names():
    global _fixture_cache
    if _fixture_cache is None:
        _fixture_cache = np.load(_FIXTURES)
    return sorted({k[:-2] for k in _fixture_cache.files})


def load_fixture(name: str, scale: float = 1.0) -> TriMesh:
    """Load one of the example meshes committed (as parsed doubles) in tests/golden/meshes.npz.

    `name` is the reference's file name without `.dat` (e.g. "G1", "s5m", "Vint16k", "Case-7-2").
    """
    global _fixture_cache
    if _fixture_cache is None:
        _fixture_cache = np.load(_FIXTURES)
    v = _fixture_cache[name + ".v"].astype(np.float64)
    c = _fixture_cache[name + ".c"].astype(np.int32)
    return TriMesh(v * float(scale), c)


def subdivide(mesh: TriMesh, levels: int, project_radius: float | None = None) -> TriMesh:
    """Uniform midpoint subdivision with shared edge vertices (synthetic meshes for the sharding sweep,
    BASELINE.json configs[4]).  Deterministic, no RNG.  Child order follows the reference's kSplitCell
    (/root/reference/src/NumericalIntegrator3d.cu:55-65)."""
    v = [tuple(p) for p in mesh.vertices.tolist()]
    cells = mesh.cells.tolist()
    for _ in range(levels):
        mid = {}
        def midpoint(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in mid:
                pa, pb = v[a], v[b]
                v.append((0.5 * (pa[0] + pb[0]), 0.5 * (pa[1] + pb[1]), 0.5 * (pa[2] + pb[2])))
                mid[key] = len(v) - 1
            return mid[key]
        new = []
        for a, b, c in cells:
            ma, mb, mc = midpoint(b, c), midpoint(c, a), midpoint(a, b)
            new += [(mc, b, ma), (ma, c, mb), (mb, a, mc), (ma, mb, mc)]
        cells = new
    verts = np.asarray(v, dtype=np.float64)
    if project_radius is not None:
        verts = verts * (project_radius / np.linalg.norm(verts, axis=1))[:, None]
    return TriMesh(verts, np.asarray(cells, dtype=np.int32))

"""Shared helpers of the parity tests: error measures and the conditioning-aware tolerance.

Tolerance statement (DESIGN.md "Parity"): per ordered pair
    |J_new - J_ref|_1  <=  1e-12 * |J_ref|_1  +  K * noise_ij ,   K = 2
where noise_ij is a first-order bound of the rounding noise of the REFERENCE's own formula
(thetaPsi, /root/reference/src/evaluators/evaluatorJ3DK.cu:266-313) for that pair.  The second term is needed
because the reference evaluates ln[l_a(1+cos)/(l_b(1+cos))] and a triple product of unit vectors: both lose
log2((l/L)^2) resp. log2(1/(1+cos)) bits for distant / nearly collinear configurations, so two compilations of
the reference itself (-fmad on/off) already differ by far more than 1e-12 there.  For well-conditioned pairs
noise_ij << 1e-12 |J| and the test is the plain 1e-12 relative bound of BASELINE.json.
"""
import numpy as np

import json
import os

U = 2.0 ** -53
K_NOISE = 2.0
REL_TOL = 1e-12

# every parity check records what it observed (how much of the tolerance was used, what fraction meets the plain 1e-12 bound):
# written by tests/conftest.py at the end of a GPU session to gpurun_out/r02_parity.json (copied to profiles/ for the record)
PARITY_LOG = {}


def record_parity(label, stats):
    if label:
        PARITY_LOG[label] = stats


def write_parity_log():
    if not PARITY_LOG:
        return None
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_dir = os.path.join(root, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "r02_parity.json")
    old = {}
    if os.path.exists(path):
        try:
            old = json.load(open(path))
        except Exception:  # noqa: BLE001
            old = {}
    old.update(PARITY_LOG)
    old["_tolerance"] = {"statement": "|J_new - J_ref|_1 <= 1e-12 |J_ref|_1 + K noise_ij", "K_NOISE_regular": K_NOISE,
                         "K_PERTURB_adjacent": K_PERTURB}
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    return path

QF13_XY = None


def _qf():
    from integrator2_b200 import abi
    xy = abi.QF13_XY
    L = np.stack([xy[:, 0], xy[:, 1], 1.0 - xy[:, 0] - xy[:, 1]], axis=1)
    return L, abi.QF13_W


def rel_err_l1(a, b):
    """|a-b|_1 / |b|_1 per row."""
    return np.abs(a - b).sum(1) / np.maximum(np.abs(b).sum(1), 1e-300)


def reference_noise_bound(vertices, cells, tasks, level=0):
    """First-order rounding-noise bound (absolute, on |J|_1) of the reference formula for regular pairs; at a refinement level
    > 0 the bound is the sum of the bounds of the 4^level children of the control panel (their own Gauss points and areas)."""
    i, j = tasks[:, 0], tasks[:, 1]
    VI = vertices[cells[i]]          # [n,3,3]
    VJ = vertices[cells[j]]
    if level > 0:
        kids = [VI]
        for _ in range(level):       # children as the reference creates them (src/NumericalIntegrator3d.cu:55-65)
            nxt = []
            for T in kids:
                A, B, C = T[:, 0], T[:, 1], T[:, 2]
                ma, mb, mc = 0.5 * (B + C), 0.5 * (C + A), 0.5 * (A + B)
                nxt += [np.stack([mc, B, ma], 1), np.stack([ma, C, mb], 1), np.stack([mb, A, mc], 1), np.stack([ma, mb, mc], 1)]
            kids = nxt
        return sum(_noise_bound_panels(T, VJ) for T in kids)
    return _noise_bound_panels(VI, VJ)


def _noise_bound_panels(VI, VJ):
    L, w = _qf()
    A, B, C = VJ[:, 0], VJ[:, 1], VJ[:, 2]
    Si = 0.5 * np.linalg.norm(np.cross(VI[:, 1] - VI[:, 0], VI[:, 2] - VI[:, 0]), axis=1)
    def unit(v):
        return v / np.linalg.norm(v, axis=1, keepdims=True)
    ta, tb, tc = unit(C - B), unit(A - C), unit(B - A)
    acc = np.zeros(VI.shape[0])
    for g in range(L.shape[0]):
        M = L[g, 0] * VI[:, 0] + L[g, 1] * VI[:, 1] + L[g, 2] * VI[:, 2]
        oa, ob, oc = unit(M - A), unit(M - B), unit(M - C)
        def d(p, q):
            return np.einsum("ij,ij->i", p, q)
        terms = 0.0
        for (o1, o2, t) in ((oa, ob, tc), (ob, oc, ta), (oc, oa, tb)):
            terms = terms + 2.0 / np.maximum(1.0 + d(o1, t), 1e-300) + 2.0 / np.maximum(1.0 + d(o2, t), 1e-300) + 4.0
        y = d(np.cross(oa, ob), oc)
        x = 1.0 + d(oa, ob) + d(ob, oc) + d(oc, oa)
        dtheta = 8.0 * (np.abs(x) + np.abs(y)) / np.maximum(x * x + y * y, 1e-300)
        acc += np.abs(w[g]) * (terms + dtheta)
    return U * acc * Si * 0.079577471545947667884


def check_regular_parity(vertices, cells, tasks, J_new, J_ref, label=""):
    """Returns a dict of statistics and asserts the tolerance statement."""
    err = np.abs(J_new - J_ref).sum(1)
    ref = np.abs(J_ref).sum(1)
    noise = reference_noise_bound(vertices, cells, tasks)
    allowed = REL_TOL * ref + K_NOISE * noise
    rel = err / np.maximum(ref, 1e-300)
    stats = dict(n=int(tasks.shape[0]), rel_median=float(np.median(rel)), rel_p99=float(np.quantile(rel, 0.99)),
                 rel_max=float(rel.max()), frac_within_1e12=float((rel <= REL_TOL).mean()),
                 worst_ratio_to_allowed=float((err / allowed).max()))
    record_parity(label, stats)
    bad = np.nonzero(err > allowed)[0]
    assert bad.size == 0, f"{label}: {bad.size} pairs outside tolerance, worst {stats}; first bad task {tasks[bad[0]]}"
    return stats


def key(tasks):
    return tasks[:, 0].astype(np.int64) * (1 << 32) + tasks[:, 1].astype(np.int64)


def align_by_pair(tasks_a, tasks_b):
    """index arrays (ia, ib) such that tasks_a[ia] and tasks_b[ib] list the same (i,j) pairs."""
    ka, kb = key(tasks_a), key(tasks_b)
    oa, ob = np.argsort(ka), np.argsort(kb)
    common, ia, ib = np.intersect1d(ka[oa], kb[ob], return_indices=True)
    return oa[ia], ob[ib]


def read_class_dump(path):
    """binary record written by oracle/ref_dump.cu: n, tasks[n][3], results[n][3], integrals[n][4]."""
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(4), dtype=np.int32)[0])
        tasks = np.frombuffer(f.read(12 * n), dtype=np.int32).reshape(n, 3).copy()
        results = np.frombuffer(f.read(24 * n), dtype=np.float64).reshape(n, 3).copy()
        integrals = np.frombuffer(f.read(32 * n), dtype=np.float64).reshape(n, 4).copy()
    return dict(tasks=tasks, results=results, integrals=integrals)


def read_mesh_dump(path):
    with open(path, "rb") as f:
        nv, nc = np.frombuffer(f.read(8), dtype=np.int32)
        v = np.frombuffer(f.read(24 * nv), dtype=np.float64).reshape(nv, 3).copy()
        c = np.frombuffer(f.read(12 * nc), dtype=np.int32).reshape(nc, 3).copy()
        nrm = np.frombuffer(f.read(24 * nc), dtype=np.float64).reshape(nc, 3).copy()
        area = np.frombuffer(f.read(8 * nc), dtype=np.float64).copy()
        adaptive = int(np.frombuffer(f.read(4), dtype=np.int32)[0])
        ref = None
        if adaptive:
            ref = [np.frombuffer(f.read(nc), dtype=np.uint8).copy() for _ in range(3)]
    return dict(vertices=v, cells=c, normals=nrm, measures=area, adaptive=adaptive, refinements=ref)


def perturbation_noise(oracle, vertices, cells, cls, tasks, level, trials=8, seed=0):
    """Empirical rounding-noise scale of the reference formula for each task: the spread of the oracle's result when
    every vertex coordinate is moved by -1/0/+1 ulp OF THE MESH EXTENT (an input change of the size of one rounding
    error of a coordinate difference, propagated through the same formulas, so it sees the same cancellations and the
    same epsilon-branches).  Used for the adjacent classes, whose regular part is the difference of two singular
    functions.  Returns (noise[n], J_base[n,3])."""
    tasks = np.ascontiguousarray(tasks)
    base = oracle.OracleMesh(vertices, cells).run_class(cls, tasks, level)["results"]
    rng = np.random.default_rng(seed)
    ulp = np.spacing(np.abs(vertices).max())
    worst = np.zeros(tasks.shape[0])
    for _ in range(trials):
        step = rng.integers(-1, 2, size=vertices.shape).astype(np.float64)
        J = oracle.OracleMesh(vertices + step * ulp, cells).run_class(cls, tasks, level)["results"]
        worst = np.maximum(worst, np.abs(J - base).sum(1))
    return worst, base


K_PERTURB = 16.0


def check_parity_perturbation(oracle, vertices, cells, cls, tasks, level, J_new, J_ref=None, label="", max_outliers=0):
    """|J_new - J_ref|_1 <= 1e-12 |J_ref|_1 + K_PERTURB * noise_ij  (J_ref defaults to the oracle's value).
    max_outliers: pairs allowed to exceed the bound (the noise estimate is itself a random sample; an epsilon-branch
    that flips between two implementations is not always reached by 8 perturbations); they must still agree to 1e-5."""
    noise, base = perturbation_noise(oracle, vertices, cells, cls, tasks, level)
    if J_ref is None:
        J_ref = base
    err = np.abs(J_new - J_ref).sum(1)
    ref = np.abs(J_ref).sum(1)
    allowed = REL_TOL * ref + K_PERTURB * noise
    rel = err / np.maximum(ref, 1e-300)
    stats = dict(n=int(tasks.shape[0]), rel_median=float(np.median(rel)), rel_p99=float(np.quantile(rel, 0.99)), rel_max=float(rel.max()),
                 frac_within_1e12=float((rel <= REL_TOL).mean()), worst_ratio_to_allowed=float((err / allowed).max()))
    bad = np.nonzero(err > allowed)[0]
    stats["outliers"] = int(bad.size)
    record_parity(label, stats)
    assert bad.size <= max_outliers, f"{label}: {bad.size} pairs outside tolerance {stats}; first bad task {tasks[bad[0]]}"
    if bad.size:
        assert rel[bad].max() < 1e-5, f"{label}: outlier too large {stats}"
    return stats


def near_singular_mesh(seed=0):
    """Synthetic mesh of disjoint triangles in crafted (i, j) couples that are REGULAR by topology but nearly singular by
    geometry: one Gauss point of i is placed next to an edge of j (distance 1e-2 .. 1e-9 edge lengths), on the line of an
    edge beyond either end (where the reference's epsilon fallback fires), next to a vertex of j, or just above j's
    interior (half solid angle near pi).  Returns (vertices, cells, couples[n][2], labels)."""
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    L3 = np.column_stack([O.QF13_XY[:, 0], O.QF13_XY[:, 1], 1.0 - O.QF13_XY[:, 0] - O.QF13_XY[:, 1]])
    verts, cells, couples, labels = [], [], [], []

    def unit(v):
        return v / np.linalg.norm(v)

    def add(P_of, label):
        m = len(couples)
        origin = np.array([8.0 * (m % 16), 8.0 * (m // 16), 0.0])
        tj = rng.normal(size=(3, 3))
        tj = np.roll(tj, m % 3, axis=0)
        P = P_of(tj)
        ti = rng.normal(size=(3, 3)) * rng.choice([0.3, 1.0])
        g = int(rng.choice([0, 1, 5, 9]))
        ti = ti + (P - L3[g] @ ti)        # Gauss point g of i lands on P
        base = len(verts)
        verts.extend((ti + origin).tolist()); verts.extend((tj + origin).tolist())
        cells.append([base, base + 1, base + 2]); cells.append([base + 3, base + 4, base + 5])
        couples.append([2 * m, 2 * m + 1]); labels.append(label)

    for delta in (1e-2, 1e-3, 1e-4, 1e-6, 1e-9, 0.0):
        for rep in range(3):
            def near_edge(tj, delta=delta):
                A, B, C = tj
                e = B - A
                n = unit(np.cross(e, rng.normal(size=3)))
                return A + rng.uniform(0.2, 0.8) * e + delta * np.linalg.norm(e) * n
            add(near_edge, f"edge interior {delta:g}")

            def beyond_b(tj, delta=delta):
                A, B, C = tj
                e = B - A
                n = unit(np.cross(e, rng.normal(size=3)))
                return A + rng.uniform(1.2, 2.0) * e + delta * np.linalg.norm(e) * n
            add(beyond_b, f"beyond B {delta:g}")

            def beyond_a(tj, delta=delta):
                A, B, C = tj
                e = B - A
                n = unit(np.cross(e, rng.normal(size=3)))
                return A - rng.uniform(0.2, 1.0) * e + delta * np.linalg.norm(e) * n
            add(beyond_a, f"beyond A {delta:g}")
    for delta in (1e-1, 1e-2, 1e-3, 1e-5):
        for rep in range(3):
            add(lambda tj, delta=delta: tj[0] + delta * np.linalg.norm(tj[1] - tj[0]) * unit(rng.normal(size=3)), f"vertex {delta:g}")
            add(lambda tj, delta=delta: tj.mean(0) + delta * unit(np.cross(tj[1] - tj[0], tj[2] - tj[0])), f"above interior {delta:g}")
            add(lambda tj, delta=delta: tj[0] + 1.7 * (tj[1] - tj[0]) + 0.9 * (tj[2] - tj[0]) + delta * 1e-3 * unit(np.cross(tj[1] - tj[0], tj[2] - tj[0])),
                f"coplanar outside {delta * 1e-3:g}")
    return np.array(verts), np.array(cells, dtype=np.int32), np.array(couples), labels


def couple_tasks(om, couples):
    """tasks of the regular class restricted to the crafted couples (both orders)."""
    t = om.tasks(2)
    keys = set(map(tuple, couples.tolist())) | set(map(tuple, couples[:, ::-1].tolist()))
    sel = np.array([(int(a), int(b)) in keys for a, b in t[:, :2]])
    return np.ascontiguousarray(t[sel])

"""List-free regular class (i2_apply_regular): row sums sum_j w_j J(K_i,K_j) over all j sharing no vertex with i, against
the oracle (explicit task list) and against the list-based kernel."""
import numpy as np
import pytest

from helpers import reference_noise_bound
from integrator2_b200.meshio import load_fixture, subdivide

pytestmark = pytest.mark.gpu


def _oracle_rowsums(om, mesh, weights, rows):
    """row sums from the explicit task list + the summed per-pair tolerance (1e-12 |J| + 8 noise_ij, helpers.py)."""
    t = om.tasks(2)
    sel = np.isin(t[:, 0], rows)
    ts = np.ascontiguousarray(t[sel])
    J = om.run_class(2, ts, 0)["results"] * weights[ts[:, 1]][:, None]
    out = np.zeros((om.n_cells, 3))
    np.add.at(out, ts[:, 0], J)
    allowed = (1e-12 * np.abs(J).sum(1) + 8.0 * reference_noise_bound(mesh.vertices, mesh.cells, ts) * weights[ts[:, 1]])
    tol = np.zeros(om.n_cells)
    np.add.at(tol, ts[:, 0], allowed)
    return out[rows], tol[rows]


@pytest.mark.parametrize("name,scale", [("G1", 1.0), ("cubehole", 1.0), ("s5m", 0.0005)])
def test_apply_regular_matches_oracle(ctx, oracle, name, scale):
    import torch
    m = load_fixture(name, scale)
    om = oracle.OracleMesh(m.vertices, m.cells)
    ctx.set_mesh(m.vertices, m.cells)
    rng = np.random.default_rng(3)
    w = rng.uniform(0.5, 1.5, m.n_cells)
    lo, hi = (0, m.n_cells) if m.n_cells < 1000 else (300, 700)
    rows = np.arange(lo, hi)
    for weights in (w, np.ones(m.n_cells)):
        got = ctx.apply_regular(lo, hi, torch.as_tensor(weights).cuda() if weights is w else None).cpu().numpy()
        ref, tol = _oracle_rowsums(om, m, weights, rows)
        err = np.abs(got - ref).sum(1)
        assert (err <= tol).all(), (name, float((err / tol).max()))
        if name == "G1":
            assert (err / np.abs(ref).sum(1)).max() < 1e-12


def test_apply_regular_equals_list_kernel_on_vint16k(ctx):
    """Same pairs, two enumerations: explicit list (i2_integrate_class) vs implicit tiles (i2_apply_regular)."""
    import torch
    m = load_fixture("Vint16k")
    ctx.set_mesh(m.vertices, m.cells)
    lists = ctx.classify()
    tasks = ctx.tasks_from_pairs(lists[2])
    J = ctx.integrate_class(2, tasks, 0, want_stats=False)["results"]
    rowsum = torch.zeros((m.n_cells, 3), dtype=torch.float64, device="cuda")
    rowsum.index_add_(0, tasks[:, 0].long(), J)
    lo, hi = 4000, 4400
    got = ctx.apply_regular(lo, hi)
    ref = rowsum[lo:hi]
    # the two kernels group lanes differently (32 consecutive tasks vs 32 columns of a tile), so their warps vote
    # differently on the far-field shortcuts and round ill-conditioned pairs differently: allow 1e-12 |J| plus the summed
    # conditioning bound of the row's pairs (helpers.reference_noise_bound), like the oracle comparison above
    t = tasks.cpu().numpy()
    sel = (t[:, 0] >= lo) & (t[:, 0] < hi)
    noise = np.zeros(m.n_cells)
    np.add.at(noise, t[sel, 0], reference_noise_bound(m.vertices, m.cells, np.ascontiguousarray(t[sel])))
    absJ = torch.zeros(m.n_cells, dtype=torch.float64, device="cuda")
    absJ.index_add_(0, tasks[:, 0].long(), J.abs().sum(1))
    tol = 1e-12 * absJ[lo:hi] + 8.0 * torch.as_tensor(noise[lo:hi]).cuda()
    err = (got - ref).abs().sum(1)
    assert bool((err <= tol).all()), float((err / tol).max())
    rel = err / ref.abs().sum(1).mean()
    assert float(rel.median()) < 1e-13 and float(rel.max()) < 1e-9


def test_apply_regular_beyond_the_reference_limit(ctx):
    """A mesh the reference cannot even load (N > 46 340: int overflow of N*N/2): G1 sphere refined 5x,
    108 544 triangles.  Property check: a closed surface seen from one of its own panels — sum_j J(K_i,K_j) is finite and
    the row sums of two symmetric panels mirror each other."""
    import torch
    g1 = load_fixture("G1")
    big = subdivide(g1, 5)
    assert big.n_cells == 106 * 4 ** 5
    ctx.set_mesh(big.vertices, big.cells)
    out = ctx.apply_regular(0, 256).cpu().numpy()
    assert np.isfinite(out).all() and np.abs(out).sum() > 0
    # children of the same parent panel lie in one plane: their row sums have the same sign pattern along the normal
    n = np.cross(big.vertices[big.cells[:256, 1]] - big.vertices[big.cells[:256, 0]], big.vertices[big.cells[:256, 2]] - big.vertices[big.cells[:256, 0]])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    along = (out * n).sum(1)
    assert (np.sign(along) == np.sign(along[0])).all()

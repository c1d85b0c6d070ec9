"""The whole operator by rows (i2_apply_*: regular class list-free + adjacent classes from row-major incidence lists) against the
list-based path on the example meshes, and against the CPU oracle on sampled rows of the two largest configurations of
BASELINE.json — s5m2 refined twice (125 280 triangles, configs[3]) and the G1 sphere refined five times (108 544 triangles,
configs[4]) — which the reference itself cannot load."""
import numpy as np
import pytest

from helpers import K_NOISE, record_parity, reference_noise_bound
from integrator2_b200.meshio import load_fixture, subdivide

pytestmark = pytest.mark.gpu


def _list_row_sums(ctx, m, level, w=None):
    """row sums of all three classes from the list-based path (i2_integrate_class), accumulated in float64 on the device"""
    import torch
    lists = ctx.classify()
    rs = torch.zeros((m.n_cells, 3), dtype=torch.float64, device="cuda")
    ab = torch.zeros(m.n_cells, dtype=torch.float64, device="cuda")
    out = []
    for cls in range(3):
        tasks = ctx.tasks_from_pairs(lists[cls])
        r = ctx.integrate_class(cls, tasks, level)
        J = r["results"] if w is None else r["results"] * w[tasks[:, 1].long()][:, None]
        rs.index_add_(0, tasks[:, 0].long(), J)
        ab.index_add_(0, tasks[:, 0].long(), J.abs().sum(1))
        out.append(r)
    return rs, ab, out


@pytest.mark.parametrize("name,scale", [("G1", 1.0), ("s5m", 0.0005), ("cubehole", 1.0)])
@pytest.mark.parametrize("level", [0, -1])
def test_apply_all_classes_equals_the_lists(ctx, name, scale, level):
    import torch
    m = load_fixture(name, scale)
    ctx.set_mesh(m.vertices, m.cells)
    w = torch.rand(m.n_cells, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) + 0.5
    for weights in (None, w):
        rs, ab, lists = _list_row_sums(ctx, m, level, weights)
        ctx.apply_prepare(0, m.n_cells)
        a = ctx.apply(level, weights)
        rel = (a["out"] - rs).abs().sum(1) / ab
        assert float(rel.median()) < 1e-13, (name, level, float(rel.median()))
        # the list-free kernel meets a pair in another warp than the list kernel does, and the far-field tier of a group's logs is
        # chosen per warp: ill-conditioned pairs see another sample of the rounding noise (the tolerance statement of helpers.py);
        # under error control a flipped tie of the list-free regular kernel changes one pair by up to ~1e-3 of its value
        assert float(torch.quantile(rel, 0.99)) < 1e-9, (name, level, float(torch.quantile(rel, 0.99)))
        assert float(rel.max()) < (2e-6 if level == 0 else 1e-4), (name, level, float(rel.max()))
        if level < 0:
            for cls in range(2):     # adjacent classes: same kernels on the same tasks -> identical counters and rounds
                assert torch.equal(a["refinements"][cls], lists[cls]["refinements"]), (name, cls)
                assert a["stats"][cls]["last_round"] == lists[cls]["stats"]["last_round"]
                assert a["stats"][cls]["unconverged"] == lists[cls]["stats"]["unconverged"]
            assert int((a["refinements"][2] != lists[2]["refinements"]).sum()) <= 8
    # a block of rows equals the same rows of the whole
    lo, hi = (m.n_cells // 3) & ~31, ((m.n_cells // 3) & ~31) + 64
    ctx.apply_prepare(0, m.n_cells)
    whole = ctx.apply(level, w)
    ctx.apply_prepare(lo, min(hi, m.n_cells))
    part = ctx.apply(level, w, split=lambda L: whole_last(whole, L))
    assert torch.equal(part["out"], whole["out"][lo:hi]), (name, level)
    assert torch.equal(part["refinements"], whole["refinements"][:, lo:hi])


def whole_last(whole, local):
    """what a multi-GPU run agrees on: the last rounds of the WHOLE job (maximum over the blocks)"""
    return [max(int(a), int(s["last_round"])) for a, s in zip(local, whole["stats"])]


def _oracle_rows(oracle, m, rows, level):
    """regular-class row sums of `rows` from the CPU oracle: (sum_j J, sum_j |J|_1, stats, refinements[rows])"""
    I = np.repeat(rows, m.n_cells)
    Jx = np.tile(np.arange(m.n_cells), rows.size)
    shared = (m.cells[I][:, :, None] == m.cells[Jx][:, None, :]).any(axis=(1, 2))
    keep = ~shared
    t = np.ascontiguousarray(np.stack([I[keep], Jx[keep], np.arange(int(keep.sum()))], axis=1).astype(np.int32))
    om = oracle.OracleMesh(m.vertices, m.cells)
    r = om.run_class(2, t, level)
    rs = np.zeros((m.n_cells, 3)); ab = np.zeros(m.n_cells)
    np.add.at(rs, t[:, 0], r["results"]); np.add.at(ab, t[:, 0], np.abs(r["results"]).sum(1))
    r["noise"] = om.noise_bound(t)      # the tolerance model of helpers.reference_noise_bound, in C/OpenMP (1e7 pairs per block)
    return rs[rows], ab[rows], r, t


@pytest.mark.parametrize("which", ["s5m2_x2_125280", "G1_x5_108544"])
def test_largest_meshes_against_the_oracle_on_sampled_rows(ctx, oracle, which):
    """>= 256 rows of each of the two largest configurations, in contiguous blocks of 64 (the value a pair ends with under error
    control depends on the last round of the set of pairs integrated together — SURVEY.md D7 — so the oracle and the device must
    see the same set: one block at a time), including the blocks whose rows refine deepest."""
    import torch
    base, scale, times = (("s5m2", 0.0005, 2) if which.startswith("s5m2") else ("G1", 1.0, 5))
    m = subdivide(load_fixture(base, scale), times)
    assert m.n_cells == int(which.rsplit("_", 1)[1])
    ctx.set_mesh(m.vertices, m.cells)
    # fixed level 0, list-free kernel: 4 blocks of 64 rows spread over the mesh
    blocks = [(int(x) & ~31) for x in np.linspace(0, m.n_cells - 64, 4)]
    worst0, within0, n0 = 0.0, [], 0
    for lo in blocks:
        rows = np.arange(lo, lo + 64)
        out = ctx.apply_regular(lo, lo + 64).cpu().numpy()
        rs, ab, r, t = _oracle_rows(oracle, m, rows, 0)
        rel = np.abs(out - rs).sum(1) / ab
        # row-level bound from the pair-level tolerance statement: sum of the allowed pair errors of the row
        allowed = np.zeros(m.n_cells)
        np.add.at(allowed, t[:, 0], 1e-12 * np.abs(r["results"]).sum(1) + K_NOISE * r["noise"])
        assert (np.abs(out - rs).sum(1) <= allowed[rows]).all(), (which, lo, float(rel.max()))
        worst0 = max(worst0, float(rel.max())); within0.append(float((rel <= 1e-12).mean())); n0 += t.shape[0]
    record_parity(f"{which} i2_apply_regular level 0: 256 sampled rows vs oracle (relative to the row's sum |J|)",
                  dict(pairs=n0, rel_max=worst0, frac_rows_within_1e12=float(np.mean(within0))))
    # error control: the whole mesh once on the device to find where it refines deepest, then 4 blocks: two evenly placed,
    # two around the rows with the highest refinement counters
    full = ctx.apply_regular_adaptive(0, m.n_cells)
    ref_all = full["refinements"].cpu().numpy()
    deepest = np.argsort(-ref_all.astype(np.int32), kind="stable")
    picks = [int(deepest[0]) & ~31]
    for d in deepest:
        if abs(int(d) - picks[0]) > 4096:
            picks.append(int(d) & ~31)
            break
    blocks = [min(b, (m.n_cells - 64) & ~31) for b in picks + [(m.n_cells // 5) & ~31, (4 * m.n_cells // 5) & ~31]]
    summary = []
    for lo in blocks:
        rows = np.arange(lo, lo + 64)
        a = ctx.apply_regular_adaptive(lo, lo + 64)
        rs, ab, r, t = _oracle_rows(oracle, m, rows, -1)
        st, so = a["stats"], r["stats"]
        L = int(so[0])
        assert st["last_round"] == L, (which, lo, st, so.tolist())
        assert st["integrated"][0] == t.shape[0]
        ties, carried = 0, 0
        for k in range(1, L + 1):
            d = abs(st["unconverged"][k] - int(so[2 + 2 * k]))
            ties = max(ties, d)
            # The list-free kernel computes a pair's round-0 and round-1 values in the same warp against the same column tile
            # (same far-field tiers), so part of their rounding noise cancels in I1 - I0; the oracle (like the reference) computes
            # them independently.  Pairs that sit on the Runge threshold only through that noise are the ones that differ:
            # observed 0.5 % of the unconverged pairs on this mesh, always towards FEWER refinements; gate at 1 % + carry.
            assert d <= max(5, 1e-2 * int(so[2 + 2 * k])) + carried, (which, lo, k, st, so.tolist())
            carried = d
        refm = a["refinements"].cpu().numpy()
        assert int((refm != r["refinements"][rows]).sum()) <= 2 * ties + 2, (which, lo)
        err = np.abs(a["out"].cpu().numpy() - rs).sum(1)
        rel = err / ab
        # row-level bound from the pair-level tolerance statement (the value of most pairs is their round-1 value: four children,
        # about the parent's noise each way -> 2 x the level-0 bound); the meshes are fine, so most pairs are far apart relative to
        # their size and the reference's formula — which the oracle restates — is ill-conditioned there: the median itself is
        # above 1e-12 on the airplane.  A flipped Runge tie changes one pair of a row by up to ~1e-3 of its value.
        allowed = np.zeros(m.n_cells)
        np.add.at(allowed, t[:, 0], 1e-12 * np.abs(r["results"]).sum(1) + 2.0 * K_NOISE * r["noise"])
        outside = err > allowed[rows]
        assert int(outside.sum()) <= 2 * ties + 2 and float(rel.max()) < 1e-4, (which, lo, int(outside.sum()), float(rel.max()), ties)
        assert float(np.median(rel)) < 1e-11, (which, lo, float(np.median(rel)))
        summary.append(dict(first_row=lo, pairs=int(t.shape[0]), last_round=L, max_counter=int(refm.max()), count_ties=ties,
                            rel_median=float(np.median(rel)), rel_max=float(rel.max()), frac_rows_within_1e12=float((rel <= 1e-12).mean()),
                            rows_outside_noise_bound=int(outside.sum()), worst_ratio_to_allowed=float((err / allowed[rows]).max())))
    assert max(s["max_counter"] for s in summary) == int(ref_all.max())
    record_parity(f"{which} i2_apply_regular_adaptive: 4 x 64 sampled rows vs oracle (relative to the row's sum |J|)",
                  dict(blocks=summary, deepest_counter_on_the_mesh=int(ref_all.max())))

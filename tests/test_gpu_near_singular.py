"""Regular-by-topology, nearly-singular-by-geometry pairs on the GPU (helpers.near_singular_mesh): exercises the rare
branches of k_regular_grouped — careful-form redo for Gauss points next to an edge / vertex of the influence triangle,
large solid angles, the reference's epsilon fallback on edge lines — at level 0, at a fixed refinement level and in the
adaptive loop, against the oracle with the conditioning-aware tolerance."""
import numpy as np
import pytest

from helpers import check_parity_perturbation, couple_tasks, near_singular_mesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [0, 1, 2])
def test_nearly_singular_regular_pairs_fixed_level(ctx, oracle, level):
    import torch
    v, c, couples, labels = near_singular_mesh()
    om = oracle.OracleMesh(v, c)
    ctx.set_mesh(v, c)
    t = couple_tasks(om, couples)
    J = ctx.integrate_class(2, torch.as_tensor(t).cuda(), level, want_stats=False)["results"].cpu().numpy()
    assert np.isfinite(J).all()
    check_parity_perturbation(oracle, v, c, 2, t, level, J, label=f"near-singular level {level}")


def test_nearly_singular_regular_pairs_whole_class_and_adaptive(ctx, oracle):
    """the complete regular class of the synthetic mesh (crafted couples + all far pairs in one launch, so warps mix
    flagged and unflagged lanes), then the adaptive loop: same rounds and counters as the oracle up to borderline ties"""
    import torch
    v, c, couples, labels = near_singular_mesh()
    om = oracle.OracleMesh(v, c)
    ctx.set_mesh(v, c)
    t = om.tasks(2)
    J = ctx.integrate_class(2, torch.as_tensor(t).cuda(), 0, want_stats=False)["results"].cpu().numpy()
    assert np.isfinite(J).all()
    check_parity_perturbation(oracle, v, c, 2, t, 0, J, label="near-singular whole class")
    r = ctx.integrate_class(2, torch.as_tensor(t).cuda(), -1)
    o = om.run_class(2, t, -1)
    L = int(o["stats"][0])
    assert r["stats"]["last_round"] == L
    for k in range(1, L + 1):
        assert abs(r["stats"]["unconverged"][k] - int(o["stats"][2 + 2 * k])) <= 3, (k, r["stats"], o["stats"].tolist())

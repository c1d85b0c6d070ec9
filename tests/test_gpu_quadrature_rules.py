"""Other Cowper rules of QuadratureFormula3d.cuh (the kernels take the rule from constant memory: any n <= 13, any run
structure of equal weights, any Runge order p)."""
import numpy as np
import pytest

from helpers import check_parity_perturbation
from integrator2_b200.meshio import load_fixture

pytestmark = pytest.mark.gpu

# (L_x, L_y) and weights as in /root/reference/src/QuadratureFormula3d.cuh: qf3D1, qf3D4, qf3D7, qf3D12
RULES = {
    "qf3D1": ([[0.3333333333333333, 0.3333333333333333]], [1.0], 2),
    "qf3D4": ([[0.3333333333333333, 0.3333333333333333], [0.6, 0.2], [0.2, 0.6], [0.2, 0.2]],
              [-0.5625, 0.5208333333333333, 0.5208333333333333, 0.5208333333333333], 3),
    "qf3D7": ([[0.3333333333333333, 0.3333333333333333], [0.101286507323456, 0.797426985353087], [0.797426985353087, 0.101286507323456],
               [0.101286507323456, 0.101286507323456], [0.470142064105115, 0.059715871789770], [0.059715871789770, 0.470142064105115],
               [0.470142064105115, 0.470142064105115]],
              [0.225, 0.125939180544827, 0.125939180544827, 0.125939180544827, 0.132394152788506, 0.132394152788506, 0.132394152788506], 5),
    "qf3D12": ([[0.873821971016996, 0.063089014491502], [0.063089014491502, 0.873821971016996], [0.063089014491502, 0.063089014491502],
                [0.501426509658179, 0.249286745170910], [0.249286745170910, 0.501426509658179], [0.249286745170910, 0.249286745170910],
                [0.636502499121399, 0.310352451033785], [0.310352451033785, 0.636502499121399], [0.636502499121399, 0.053145049844816],
                [0.053145049844816, 0.636502499121399], [0.310352451033785, 0.053145049844816], [0.053145049844816, 0.310352451033785]],
               [0.050844906370207] * 3 + [0.116786275726379] * 3 + [0.082851075618374] * 6, 6),
}


@pytest.mark.parametrize("rule", sorted(RULES))
def test_other_quadrature_rules(oracle, rule):
    import torch
    from integrator2_b200 import abi
    xy, w, order = RULES[rule]
    ctx = abi.Context(0)
    try:
        ctx.set_quadrature(np.array(xy), np.array(w), order)
        oracle.set_quadrature(np.array(xy), np.array(w), order)
        m = load_fixture("G1")
        om = oracle.OracleMesh(m.vertices, m.cells)
        ctx.set_mesh(m.vertices, m.cells)
        for cls in range(3):
            tasks = om.tasks(cls)
            for level in (0, 1):
                r = ctx.integrate_class(cls, torch.as_tensor(tasks).cuda(), level)
                ref = om.run_class(cls, tasks, level)
                rel = np.abs(r["results"].cpu().numpy() - ref["results"]).sum(1) / np.abs(ref["results"]).sum(1)
                assert rel.max() < 2e-12, (rule, cls, level, rel.max())
            a = ctx.integrate_class(cls, torch.as_tensor(tasks).cuda(), -1)
            ra = om.run_class(cls, tasks, -1)
            assert a["stats"]["last_round"] == int(ra["stats"][0])
            for k in range(1, int(ra["stats"][0]) + 1):
                assert abs(a["stats"]["unconverged"][k] - int(ra["stats"][2 + 2 * k])) <= max(2, 0.02 * int(ra["stats"][2 + 2 * k])), (rule, cls, k)
    finally:
        ctx.close()
        oracle.set_quadrature(oracle.QF13_XY, oracle.QF13_W, oracle.QF13_ORDER)
        c2 = abi.Context(0)   # restore the process-global constant memory to the 13-point rule
        c2.close()

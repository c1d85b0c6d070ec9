"""Error behaviour of the C ABI on a device: negative I2_E_* codes for misuse, 0 for empty work, messages for every code."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_error_codes_and_empty_inputs():
    import torch
    from integrator2_b200 import abi
    from integrator2_b200.meshio import load_fixture
    L = abi.load_library()
    h = C.c_void_p()
    assert L.i2_create(C.byref(h), 0) == 0
    t = torch.zeros((4, 3), dtype=torch.int32, device="cuda")
    I = torch.zeros((4, 4), dtype=torch.float64, device="cuda")
    J = torch.zeros((4, 3), dtype=torch.float64, device="cuda")
    p = lambda x: C.c_void_p(x.data_ptr())
    # no mesh yet
    assert L.i2_integrate_class(h, 2, p(t), 4, 0, p(I), p(J), None, None, None) == -2      # I2_E_NOMESH
    assert L.i2_apply_regular(h, 0, 1, None, p(J)) == -2
    m = load_fixture("G1")
    ctx = abi.Context(0)
    ctx.set_mesh(m.vertices, m.cells)
    # misuse on a prepared context
    assert L.i2_integrate_class(ctx.h, 3, p(t), 4, 0, p(I), p(J), None, None, None) == -1  # class out of range
    assert L.i2_integrate_class(ctx.h, 2, p(t), -1, 0, p(I), p(J), None, None, None) == -1
    assert L.i2_integrate_class(ctx.h, 2, None, 4, 0, p(I), p(J), None, None, None) == -1
    assert L.i2_integrate_class(ctx.h, 2, p(t), 4, 13, p(I), p(J), None, None, None) == -4  # I2_E_LEVEL
    assert L.i2_apply_regular(ctx.h, 0, m.n_cells + 1, None, p(J)) == -1
    assert L.i2_set_math_mode(ctx.h, 9) == -1
    # empty work is not an error
    st = abi.Stats()
    assert L.i2_integrate_class(ctx.h, 0, None, 0, -1, None, None, None, None, C.byref(st)) == 0
    assert st.last_round == 0
    assert L.i2_symmetry_error(ctx.h, None, 0, None) == 0
    for code in (0, -1, -2, -3, -4, -5, 2, 700):
        assert L.i2_error_string(code)
    # quadrature not set on the bare context -> I2_E_NOQUAD after a mesh is attached
    assert L.i2_set_mesh(h, p(ctx.d_vertices), 55, p(ctx.d_cells), m.n_cells, p(ctx.d_normals), p(ctx.d_measures)) == 0
    assert L.i2_integrate_class(h, 2, p(t), 4, 0, p(I), p(J), None, None, None) == -3      # I2_E_NOQUAD
    ctx.close()
    assert L.i2_destroy(h) == 0


def test_two_contexts_share_the_process_wide_quadrature(oracle):
    """Like the reference's __constant__ symbols, the Gauss rule is process-global: a second context sees the same rule."""
    import numpy as np
    import torch
    from integrator2_b200 import abi
    from integrator2_b200.meshio import load_fixture
    m = load_fixture("G1")
    a, b = abi.Context(0), abi.Context(0)
    a.set_mesh(m.vertices, m.cells)
    b.set_mesh(m.vertices, m.cells)
    om = oracle.OracleMesh(m.vertices, m.cells)
    t = torch.as_tensor(om.tasks(1)).cuda()
    ra = a.integrate_class(1, t, 1)["results"]
    rb = b.integrate_class(1, t, 1)["results"]
    assert torch.equal(ra, rb)                                   # deterministic, bit-identical between contexts
    a.close()
    b.close()


def test_results_are_bitwise_reproducible(ctx, oracle):
    """No atomics on the value path (the reference sums refined results with FP64 atomicAdd in arbitrary order): two runs of
    the same call give identical bits, in fixed and in adaptive mode."""
    import torch
    from integrator2_b200.meshio import load_fixture
    m = load_fixture("s5m", 0.0005)
    ctx.set_mesh(m.vertices, m.cells)
    om = oracle.OracleMesh(m.vertices, m.cells)
    for cls in (0, 2):
        t = torch.as_tensor(om.tasks(cls)).cuda()
        for level in (2, -1):
            r1 = ctx.integrate_class(cls, t, level)
            r2 = ctx.integrate_class(cls, t, level)
            assert torch.equal(r1["results"], r2["results"]) and torch.equal(r1["integrals"], r2["integrals"])
            if level < 0:
                assert r1["stats"] == r2["stats"] and torch.equal(r1["refinements"], r2["refinements"])

"""Accuracy of the branch-free FP64 primitives of the regular-pair kernel (integrator2_b200/csrc/i2_math.cuh).

CPU: the header is compiled as host code (tests/host_emu) with the hardware seeds emulated at 2^-20 accuracy (worse
than MUFU.RCP64H / RSQ64H), which checks the algorithms.  GPU: the same functions run on the device through the
i2_selftest_math hook.  Bar: <= 2 ulp against a long-double (64-bit mantissa) evaluation."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

LD = np.longdouble


def _inputs():
    rng = np.random.default_rng(1)
    n = 400_000
    x = np.exp(rng.uniform(-60, 60, n))
    a = np.exp(rng.uniform(-30, 30, n))
    b_wide = a * np.exp(rng.uniform(-3, 3, n))
    b_near = a * (1 + rng.uniform(-1e-3, 1e-3, n))
    y = rng.normal(size=n) * np.exp(rng.uniform(-100, 100, n))
    xx = rng.normal(size=n) * np.exp(rng.uniform(-100, 100, n))
    return x, a, b_wide, b_near, y, xx


def _ulps(got, ref):
    ref64 = ref.astype(np.float64)
    return np.abs(got.astype(LD) - ref) / np.spacing(np.abs(ref64)).astype(LD)


def _check(fn_sqrt, fn_rcp, fn_log, fn_atan2):
    x, a, b_wide, b_near, y, xx = _inputs()
    assert _ulps(fn_sqrt(x), np.sqrt(x.astype(LD))).max() <= 2
    assert _ulps(fn_rcp(x), 1 / x.astype(LD)).max() <= 2
    # wide ratios: |log| is O(1); absolute accuracy vs log(a) - log(b) in extended precision
    ref = np.log(a.astype(LD)) - np.log(b_wide.astype(LD))
    got = fn_log(a, b_wide)
    big = np.abs(ref) > 0.05
    assert _ulps(got[big], ref[big]).max() <= 3
    assert np.abs(got.astype(LD) - ref).max() < 1e-15
    # ratios near 1: full RELATIVE accuracy (where log(a/b) loses all of it)
    ref = np.log1p((a.astype(LD) - b_near.astype(LD)) / b_near.astype(LD))
    assert _ulps(fn_log(a, b_near), ref).max() <= 3
    ref = np.arctan2(y.astype(LD), xx.astype(LD))
    assert _ulps(fn_atan2(y, xx), ref).max() <= 3
    ys = np.array([0.0, 1.0, -1.0, 0.0, 1e-300, -0.0, 1.0, 1.0, 3e200, -2e-200])
    xs = np.array([1.0, 0.0, 0.0, -1.0, 1.0, -1.0, 1.0, -1.0, 1e-200, -4e200])
    assert np.allclose(fn_atan2(ys, xs), np.arctan2(ys, xs), rtol=1e-15, atol=0)


def test_primitives_host_emulation():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "host_emu")], check=True)
    emu = C.CDLL(os.path.join(ROOT, "tests", "host_emu", "libemu.so"))
    dp = C.POINTER(C.c_double)

    def one(fn):
        def f(x):
            x = np.ascontiguousarray(x)
            o = np.empty_like(x)
            fn(x.ctypes.data_as(dp), C.c_longlong(x.size), o.ctypes.data_as(dp))
            return o
        return f

    def two(fn):
        def f(a, b):
            a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
            o = np.empty_like(a)
            fn(a.ctypes.data_as(dp), b.ctypes.data_as(dp), C.c_longlong(a.size), o.ctypes.data_as(dp))
            return o
        return f

    _check(one(emu.emu_fast_sqrt), one(emu.emu_fast_rcp), two(emu.emu_log_ratio), two(emu.emu_atan2))


@pytest.mark.gpu
def test_primitives_on_device(ctx):
    import torch

    def one(op):
        return lambda x: ctx.selftest_math(op, torch.as_tensor(np.ascontiguousarray(x)).cuda()).cpu().numpy()

    def two(op):
        return lambda a, b: ctx.selftest_math(op, torch.as_tensor(np.ascontiguousarray(a)).cuda(), torch.as_tensor(np.ascontiguousarray(b)).cuda()).cpu().numpy()

    _check(one(0), one(1), two(2), two(3))

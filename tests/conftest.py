import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def _ensure_built():
    """Build steps (not fallbacks): compile the native pieces if a fresh checkout lacks them (nvcc cross-compiles without a GPU)."""
    import shutil
    import subprocess
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    targets = [("integrator2_b200/libintegrator2_b200.so", ["make", "-C", "integrator2_b200/csrc"]),
               ("integrator2_b200/host/integrator2test3D", ["make", "-C", "integrator2_b200/host"]),
               ("oracle/liboracle.so", ["make", "-C", "oracle", "liboracle.so"]),
               ("tests/compat/ref_compat_dump", ["make", "-C", "tests/compat"])]
    for out, cmd in targets:
        if not os.path.exists(os.path.join(ROOT, out)) and (shutil.which("nvcc") or "oracle" in out):
            subprocess.run(cmd, cwd=ROOT, env=env, check=False, capture_output=True)


_ensure_built()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def ctx():
    """One integrator context on cuda:0 for the whole GPU session (fails loudly without the CUDA library)."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from integrator2_b200 import abi
    c = abi.Context(0)
    yield c
    c.close()


def pytest_sessionfinish(session, exitstatus):
    """GPU sessions leave a record of how much of the parity tolerance was actually used (gpurun_out/r02_parity.json)."""
    try:
        import helpers
        if helpers.PARITY_LOG and any("gpu" in str(getattr(i, "keywords", {})) for i in session.items[:1] or [None]) is not None:
            import torch
            if torch.cuda.is_available():
                helpers.write_parity_log()
    except Exception:  # noqa: BLE001
        pass

"""integrator2_b200 — B200-native (sm_100a) implementation of integrator2's hot path.

The product is the CUDA shared library `libintegrator2_b200.so` (kernels + C ABI, include/i2_abi.h) and the
drop-in C++ host classes / CLI built on top of it (include/integrator2/, integrator2_b200/host/).  This Python
package only carries the ctypes binding used by tests/ and bench.py and the mesh readers.
"""
from . import meshio  # noqa: F401

__all__ = ["meshio", "abi"]

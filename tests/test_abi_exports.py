"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/i2_abi.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "i2_abi.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char \*|i2_context \*)\s*\*?\s*(i2_[a-z0-9_]+)\s*\(", text, flags=re.M)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ("i2_create", "i2_set_quadrature", "i2_set_mesh", "i2_integrate_class", "i2_symmetry_error",
                 "i2_classify_count", "i2_classify_fill", "i2_host_prepare", "i2_host_run"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from integrator2_b200 import abi
    lib = ctypes.CDLL(abi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/i2_abi.h but not exported"
    assert set(abi.EXPORTS) == set(_declared())


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product must fail loudly (a CUDA error code), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    from integrator2_b200 import abi
    L = abi.load_library()
    h = ctypes.c_void_p()
    rc = L.i2_create(ctypes.byref(h), 0)
    assert rc > 0, "i2_create must return a cudaError_t when no device is present"
    assert L.i2_error_string(rc)


def test_product_never_links_or_imports_the_oracle():
    """The oracle is test infrastructure: nothing under integrator2_b200/ or include/ may reference it."""
    bad = []
    for base in ("integrator2_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or fn == "Makefile":
                    txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                    if re.search(r"liboracle|oracle_py|from oracle|import oracle|orc_", txt):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_ctypes_binding_matches_the_header_prototypes():
    """Every prototype of include/i2_abi.h has the same number of parameters as the argtypes the ctypes binding declares
    (a mismatch would corrupt the call on the GPU box, where it is expensive to find)."""
    from integrator2_b200 import abi
    L = abi.load_library()
    text = open(os.path.join(ROOT, "include", "i2_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"(?:int|const char \*|i2_context \*)\s*\*?\s*(i2_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) == len(_declared())
    for name, params in protos:
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(L, name)
        assert fn.argtypes is not None, f"{name}: no argtypes in abi.py"
        assert len(fn.argtypes) == n, f"{name}: header has {n} parameters, abi.py declares {len(fn.argtypes)}"


def test_header_is_plain_c():
    """The boundary is a C ABI: include/i2_abi.h must compile as C99 (no C++-only constructs, no torch or CUDA types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    r = subprocess.run([gcc, "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-pedantic", os.path.join(ROOT, "include", "i2_abi.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "i2_abi.h")).read(), flags=re.S)     # comments may mention torch
    assert "torch" not in code.lower() and "cudaStream_t" not in code and "#include" not in code

"""The C-ABI library builds, loads and exports every symbol declared in include/nmgp_b200.h (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nmgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nmgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    from collaborative_nonstationary_multivariate_gaussian_process_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.nmgp_version.restype = ctypes.c_int
    assert lib.nmgp_version() >= 100


def test_every_wrapper_binds_a_declared_symbol():
    """_ops.py may only call entry points that the header declares."""
    src = open(os.path.join(ROOT, "collaborative_nonstationary_multivariate_gaussian_process_b200", "_ops.py")).read()
    used = set(re.findall(r"lib\(\)\.(nmgp_[a-z0-9_]+)", src))
    assert used <= set(declared_symbols()), used - set(declared_symbols())


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "collaborative_nonstationary_multivariate_gaussian_process_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn

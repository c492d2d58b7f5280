"""ctypes binding of the C-ABI shared library ``csrc/libnmgp_b200.so``
(declared in include/nmgp_b200.h).  There is no fallback: if the library is
missing the first kernel call raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnmgp_b200.so")
_lib = None


class NMGPLibraryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NMGPLibraryError(
                "CUDA extension %s not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "this package has no CPU fallback" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.nmgp_last_error.restype = ctypes.c_char_p
        _lib.nmgp_version.restype = ctypes.c_int
    return _lib


def check(status: int, what: str) -> None:
    """0 ok; <0 argument error; >0 numerical failure (e.g. 1+index of a non-PD matrix).
    Raised as RuntimeError, the class torch.cholesky/torch.solve raise in the reference."""
    if status != 0:
        msg = lib().nmgp_last_error()
        raise RuntimeError("%s failed (status %d): %s" % (what, status, msg.decode() if msg else ""))

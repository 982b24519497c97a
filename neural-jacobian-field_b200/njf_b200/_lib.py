"""ctypes binding of ``libnjf_b200.so`` (the C-ABI declared in ``include/njf_b200.h``).

The library is plain CUDA (no torch types in any signature); PyTorch is only used by the
callers for device memory and streams.  There is deliberately NO fallback: if the shared
library is missing, or a call returns non-zero, we raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NJF_LIB") or os.path.normpath(os.path.join(_HERE, "..", "lib", "libnjf_b200.so"))

_lib = None


class NjfError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NjfError(
                f"{LIB_PATH} not found - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the render path)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().njf_last_error()
        raise NjfError(msg.decode() if msg else f"libnjf_b200 call failed with status {status}")


def _declare(L: ctypes.CDLL) -> None:
    L.njf_last_error.restype = c_char_p
    L.njf_last_error.argtypes = []
    L.njf_version.restype = c_int
    L.njf_version.argtypes = []
    L.njf_selftest_chain.restype = c_int
    L.njf_selftest_chain.argtypes = [c_void_p] * 9 + [c_int, c_int, c_void_p]
    for name, (res, args) in _OPTIONAL.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args


# filled in by api.py once the render entry points exist (kept separate so that the
# self-test can run against a library that only contains the chain machinery)
_OPTIONAL: dict = {}


def ptr(t) -> c_void_p:
    """Device/host data pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return c_void_p(0)
    assert t.is_contiguous(), "libnjf_b200 takes contiguous buffers"
    return c_void_p(t.data_ptr())

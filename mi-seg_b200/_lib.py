"""ctypes binding of the C ABI in include/micn.h (libmicn.so, built in-tree by `__graft_entry__.build()`).

The shared library IS the product path: there is no CPU or PyTorch fallback.  If the library is
missing or does not export every symbol the header declares, importing this module's `lib()` raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmicn.so")

# dtype / epilogue codes of include/micn.h
MICN_F32, MICN_BF16, MICN_F16 = 0, 1, 2
EPI_NONE, EPI_LRELU, EPI_ADD_LRELU, EPI_NORM_ADD_LRELU = 0, 1, 2, 3
ERR_UNSUPPORTED = -7
MAX_STYLES = 16
MAX_PEERS = 16
FOLD_NONE, FOLD_THIS, FOLD_PREVIOUS = 0, 1, 2

c_void_p, c_int, c_int64, c_size_t, c_float = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t,
                                               ctypes.c_float)

# every symbol include/micn.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "micn_version": (c_int, []),
    "micn_error_string": (ctypes.c_char_p, [c_int]),
    "micn_set_option": (c_int, [ctypes.c_char_p, ctypes.c_longlong]),
    "micn_get_option": (ctypes.c_longlong, [ctypes.c_char_p]),
    "micn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "micn_read_status": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_int)]),
    "micn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                         c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_float,
                         c_void_p, c_size_t, c_void_p]),
    "micn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                         c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "micn_peer_buffer_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "micn_bwd_allreduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                   c_int, c_int, c_float, c_void_p, c_size_t, c_void_p, c_int, c_int, c_int, c_void_p]),
    "micn_allreduce_fold": (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "micn_fwd_prelu": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                               c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_float,
                               c_void_p, c_size_t, c_void_p]),
    "micn_bwd_prelu": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                               c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "micn_dual_supported": (c_int, [c_int64, c_int64, c_int64, c_int, c_int]),
    "micn_fwd_dual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_float, c_float,
                              c_void_p, c_size_t, c_void_p]),
    "micn_bwd_dual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int64, c_int64, c_int64, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "micn_cl_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "micn_fwd_cl": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                            c_int64, c_int64, c_int64, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "micn_bwd_cl": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "micn_host_scratch_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int, c_int]),
    "micn_fwd_bwd_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_float,
                                  c_void_p, c_size_t]),
}

_lock = threading.Lock()
_lib = None


class MicnError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load libmicn.so once; raise loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MicnError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C mi-seg_b200/csrc`).  There is no CPU / PyTorch fallback for instance_cond.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().micn_error_string(rc)
        raise MicnError(f"{what} failed: rc={rc} ({msg.decode() if msg else '?'})")


option_generation = 0  # bumped by every set_option: plans cached on the Python side (dual_supported) key on it


def set_option(key: str, value: int) -> None:
    global option_generation
    check(lib().micn_set_option(key.encode(), int(value)), f"micn_set_option({key})")
    option_generation += 1


def get_option(key: str) -> int:
    return int(lib().micn_get_option(key.encode()))

"""ctypes binding of the C ABI in include/dycon_b200.h.

There is deliberately no fallback: if ``_dycon_b200.so`` is missing or the current
device is not a B200-class (sm_100) GPU, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# DYCON_SO_VARIANT=timeline loads the measurement build (DYCON_TIMELINE=1 python -m ...csrc.build; tools/timeline.py)
_VARIANT = os.environ.get("DYCON_SO_VARIANT", "")
SO_PATH = os.path.join(_HERE, f"_dycon_b200_{_VARIANT}.so" if _VARIANT else "_dycon_b200.so")

FECL_FP32 = 0
FECL_BF16 = 1
FECL_FP16 = 2
ABI_VERSION = 2
EXCHANGE_NONE, EXCHANGE_UNCL, EXCHANGE_FECL, EXCHANGE_FECL_TEACHER = 0, 1, 2, 3
EXCHANGE_CHANNELS = 3
LABEL_INT64, LABEL_FLOAT32, LABEL_UINT8 = 0, 1, 2

_lock = threading.Lock()
_lib = None

_f = C.c_float
_d = C.c_double
_i = C.c_int
_i64 = C.c_int64
_p = C.c_void_p
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/dycon_b200.h one to one
PROTOTYPES = {
    "dycon_abi_version": (_i, []),
    "dycon_last_error": (C.c_char_p, []),
    "dycon_device_check": (_i, []),
    "dycon_launch_count": (C.c_uint64, []),
    "dycon_uncl_workspace_bytes": (_sz, []),
    "dycon_uncl_fwd": (_i, [_p, _p, _i64, _i, _i64, _f, _d, _p, _p, _p, _p, _sz, _p]),
    "dycon_uncl_bwd": (_i, [_p, _p, _p, _i64, _i, _i64, _f, _d, _p, _p, _p]),
    "dycon_fecl_state_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dycon_fecl_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "dycon_fecl_fwd": (_i, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _p, _i, _i, _i,
                            _f, _f, _i, _f, _f, _d, _i, _p, _sz, _p, _p, _p, _sz, _p]),
    "dycon_fecl_bwd": (_i, [_p, _sz, _p, _i, _i, _i, _i, _f, _f, _i, _i, _f, _f, _i, _p, _p, _p, _i64, _i64, _i64, _p]),
    "dycon_debug_timeline": (_sz, [_p, _sz]),
    "dycon_fecl_gn_state_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dycon_fecl_gn_layout": (_i, [_i, _i, _i, _i, _i, _p]),
    "dycon_fecl_gn_fwd": (_i, [_i, _p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _p, _i, _i, _i, _f, _f, _i, _f, _f,
                               _i, _p, _sz, _i, _i, _p, _p, _sz, _p]),
    "dycon_fecl_gn_bwd": (_i, [_p, _sz, _p, _i, _i, _i, _i, _f, _f, _i, _i, _f, _f, _i, _i, _i, _p, _p, _p,
                               _i64, _i64, _i64, _p]),
    "dycon_segcons_workspace_bytes": (_sz, []),
    "dycon_segcons_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i64, _f, _p, _p, _p, _sz, _p]),
    "dycon_segcons_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i64, _f, _p, _p, _p, _p]),
    "dycon_pool_mask": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "dycon_row_inv_norm": (_i, [_p, _i64, _i64, _i64, _i, _i, _i, _p, _p]),
    "dycon_normalize_bwd": (_i, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _i, _i, _i, _p, _i64, _i64, _i64, _p]),
    "dycon_fecl_fwd_scaled": (_i, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _p, _p, _p, _i, _i, _i,
                                   _f, _f, _i, _f, _f, _d, _i, _p, _sz, _p, _p, _p, _sz, _p, _i, _i, _p, _d, _p]),
    "dycon_grad_norm_workspace_bytes": (_sz, []),
    "dycon_grad_norm": (_i, [_p, _p, _i, _f, _p, _p, _sz, _p]),
    "dycon_sgd_ema_step": (_i, [_p, _p, _p, _p, _p, _i, _f, _f, _f, _i, _i, _f, _f, _p, _p, _i, _p]),
    "dycon_finite_check": (_i, [_p, _i, _p, _p, _p]),
    "dycon_ema_multi": (_i, [_p, _p, _p, _i, _f, _f, _p]),
    "dycon_exchange_inbox_bytes": (_sz, []),
    "dycon_exchange_enable_peer": (_i, [_i]),
    "dycon_exchange_error_offset": (_sz, []),
    "dycon_exchange_sums": (_i, [_p, _i, _p, _p, _i, _i, _p, _i, _d, _d, _p, _d, _p]),
    "dycon_uncl_fwd_sharded": (_i, [_p, _p, _i64, _i, _i64, _f, _d, _p, _p, _p, _p, _sz, _p, _i, _i, _p, _d, _p]),
    "dycon_fecl_fwd_sharded": (_i, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _p, _i, _i, _i,
                                    _f, _f, _i, _f, _f, _d, _i, _p, _sz, _p, _p, _p, _sz, _p, _i, _i, _p, _d, _p]),
}


class DyconError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(SO_PATH):
                raise DyconError(
                    f"{SO_PATH} is missing: build it with `python -m dycon_paper_replication_b200.csrc.build` "
                    "(there is no CPU or PyTorch fallback for the DyCON loss kernels)")
            handle = C.CDLL(SO_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(handle, name)      # AttributeError if the ABI and the header diverge
                fn.restype = res
                fn.argtypes = args
            if handle.dycon_abi_version() != ABI_VERSION:
                raise DyconError("dycon ABI version mismatch")
            _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dycon_last_error().decode("utf-8", "replace")
        raise DyconError(f"{what}: {msg} (code {rc})")


_device_ok = set()


def require_b200(device_index: int):
    """Raise unless the *current* CUDA device (already set by the caller) is sm_100."""
    if device_index in _device_ok:
        return
    check(lib().dycon_device_check(), "dycon_device_check")
    _device_ok.add(device_index)

"""The C-ABI library loads on a CPU-only box and exports exactly what include/dycon_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "dycon_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|uint64_t|const char\*)\s+(dycon_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


@pytest.fixture(scope="module")
def so_path():
    from dycon_paper_replication_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        from dycon_paper_replication_b200.csrc.build import build
        build()
    return _lib.SO_PATH


def test_header_parses():
    fns = header_functions()
    assert {"dycon_uncl_fwd", "dycon_uncl_bwd", "dycon_fecl_fwd", "dycon_fecl_bwd", "dycon_ema_multi",
            "dycon_last_error"} <= set(fns)


def test_library_exports_every_declared_symbol(so_path):
    handle = ctypes.CDLL(so_path)
    for name in header_functions():
        assert hasattr(handle, name), f"{name} declared in dycon_b200.h but not exported"


def test_ctypes_prototypes_match_header(so_path):
    from dycon_paper_replication_b200 import _lib
    fns = header_functions()
    assert set(_lib.PROTOTYPES) == set(fns)
    for name, nargs in fns.items():
        assert len(_lib.PROTOTYPES[name][1]) == nargs, name


def test_size_queries_without_gpu(so_path):
    from dycon_paper_replication_b200 import _lib
    L = _lib.lib()
    assert L.dycon_abi_version() == _lib.ABI_VERSION
    assert L.dycon_uncl_workspace_bytes() >= 16
    # fp32 state: student + teacher operand copies + 4 stat planes
    b, n, d = 4, 1728, 256
    assert L.dycon_fecl_state_bytes(b, n, d, 1, _lib.FECL_FP32) >= 2 * b * n * d * 4 + 4 * b * n * 4
    assert L.dycon_fecl_state_bytes(b, n, d, 0, _lib.FECL_FP32) < L.dycon_fecl_state_bytes(b, n, d, 1, _lib.FECL_FP32)
    assert L.dycon_fecl_state_bytes(0, n, d, 0, _lib.FECL_FP32) == 0


def test_argument_validation_happens_before_any_launch(so_path):
    """Rejected calls return a negative code and a message; nothing touches the (absent) GPU."""
    from dycon_paper_replication_b200 import _lib
    L = _lib.lib()
    rc = L.dycon_uncl_fwd(None, None, 1, 2, 8, 1.0, 1.0, None, None, None, None, 0, None)
    assert rc == -1 and b"NULL" in L.dycon_last_error()
    rc = L.dycon_uncl_fwd(None, None, 0, 2, 8, 1.0, 1.0, None, None, None, None, 0, None)
    assert rc == -1
    rc = L.dycon_fecl_bwd(None, 0, None, 1, 8, 8, 0, 1.0, 2.0, 0, 0, 0.3, 1.0, 7, None, None, None, 64, 8, 1, None)
    assert rc == -1 and b"precision" in L.dycon_last_error()
    # the tensor-core kernels rely on cross_thresh >= 0 (zero padding must not look like a hard negative)
    import ctypes
    buf = (ctypes.c_char * 4096)()
    addr = (ctypes.addressof(buf) + 255) // 256 * 256
    ptr = ctypes.c_void_p(addr)
    big = 1 << 30
    rc = L.dycon_fecl_fwd(ptr, 64, 8, 1, ptr, 64, 8, 1, ptr, None, 1, 8, 8, 1.0, 2.0, 1, -0.1, 1.0, 0.125,
                          _lib.FECL_FP16, ptr, big, ptr, None, ptr, big, None)
    assert rc == -2 and b"cross_thresh" in L.dycon_last_error()
    assert L.dycon_ema_multi(None, None, None, 0, 0.5, 0.5, None) == 0      # empty list is a no-op
    assert L.dycon_ema_multi(None, None, None, 3, 0.5, 0.5, None) == -1

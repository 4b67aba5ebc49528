"""Build dycon_paper_replication_b200/_dycon_b200.so with nvcc for sm_100a (in-tree).

    python -m dycon_paper_replication_b200.csrc.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box
with the gpurun snapshot.  A content hash of the sources is stored beside it so that
``build()`` is a no-op when nothing changed.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
# DYCON_TIMELINE=1 builds the measurement variant beside the product library (selected at run time with
# DYCON_SO_VARIANT=timeline, see _lib.py)
# DYCON_VARIANT=name DYCON_VARIANT_FLAGS="-DX=1 ..." builds an experiment variant _dycon_b200_<name>.so the same way
_VARIANT = "timeline" if os.environ.get("DYCON_TIMELINE") == "1" else os.environ.get("DYCON_VARIANT", "")
SO = os.path.join(PKG, f"_dycon_b200_{_VARIANT}.so" if _VARIANT else "_dycon_b200.so")
STAMP = SO + ".hash"
SOURCES = ["api.cu", "uncl.cu", "segcons.cu", "prep.cu", "ema.cu", "sgd_ema.cu", "exchange.cu", "fecl_api.cu", "fecl_simt.cu", "fecl_tc.cu", "tc_host.cu"]
HEADERS = ["common.cuh", "exchange.cuh", "fecl_math.cuh", "fecl_internal.h", "tc_common.cuh", os.path.join(ROOT, "include", "dycon_b200.h")]
# DYCON_TIMELINE=1: a measurement build whose FeCL kernels record device-clock stamps (tools/timeline.py)
_EXTRA = ["-DDYCON_TIMELINE"] if os.environ.get("DYCON_TIMELINE") == "1" else os.environ.get("DYCON_VARIANT_FLAGS", "").split()
NVCC_FLAGS = _EXTRA + [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--use_fast_math" if False else "-DDYCON_NO_GLOBAL_FAST_MATH",   # fast intrinsics are chosen per call site
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
    "-I", os.path.join(ROOT, "include"), "-I", HERE,
]


def _nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built on this machine")
    return path


def _digest(sources):
    h = hashlib.sha256()
    for f in sources + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(HERE, f)
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    sources = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    digest = _digest(sources)
    if not force and os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return SO
    objs = []
    build_dir = os.path.join(ROOT, "build", f"obj_{_VARIANT}" if _VARIANT else "obj")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for s in sources:
        obj = os.path.join(build_dir, s.replace(".cu", ".o"))
        cmd = [_nvcc(), "-c", os.path.join(HERE, s), "-o", obj] + [f for f in NVCC_FLAGS if f != "-shared"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out, file=sys.stderr if p.returncode else sys.stdout)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    link = [_nvcc(), "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(link, check=True)
    open(STAMP, "w").write(digest)
    return SO


def build_probe():
    """tests/native/tc_probe.cu -> build/tc_probe (stand-alone tcgen05 / TMA encoding check)."""
    out = os.path.join(ROOT, "build", "tc_probe")
    src = [os.path.join(ROOT, "tests", "native", "tc_probe.cu"), os.path.join(HERE, "tc_host.cu"),
           os.path.join(HERE, "api.cu")]
    newest = max(os.path.getmtime(f) for f in src + [os.path.join(HERE, "tc_common.cuh")])
    if os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-lineinfo", "-std=c++17",
           "-I", os.path.join(ROOT, "include"), "-I", HERE] + src + ["-o", out]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_probe())

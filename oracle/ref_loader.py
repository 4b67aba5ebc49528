"""Locate and import the UNMODIFIED reference modules -- TEST INFRASTRUCTURE, NOT PRODUCT.

The reference (rogeliorjr/DyCON_Paper_Replication) is pure Python.  In the build container it lives at
``/root/reference``; the GPU box has no such path, so ``stage()`` (called by ``__graft_entry__.build()``
here) copies the few files the harnesses execute into the git-ignored ``baseline/_ref/`` -- which travels
with the gpurun snapshot but never enters the history.  Nothing under ``baseline/_ref`` is product source and
the product package never imports it: only ``bench.py --impl reference`` (the reference arm, kind
"reference"), ``bench.py --train-step`` (config 3: the reference UNet3D as context) and the tests use it.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = os.environ.get("DYCON_REFERENCE", "/root/reference")
STAGED = os.path.join(ROOT, "baseline", "_ref")
# the loss module (the hot path itself), the stock losses the step loop adds to it (f2's oracle) and the
# UNet3D used by every run script (config 3's context model) with its two local imports
FILES = ["code/utils/dycon_losses.py", "code/utils/losses.py", "code/utils/ramps.py",
         "code/networks/UNet3D_contrastive.py", "code/networks/utils.py", "code/networks/networks_other.py",
         "code/networks/assp.py"]


def stage() -> bool:
    """Copy FILES from /root/reference to baseline/_ref (byte-identical).  Returns False without a reference."""
    if not os.path.isdir(SOURCE):
        return False
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(STAGED, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
    return True


def root():
    """Directory that holds the reference's ``code/`` tree: /root/reference first, then the staged copy."""
    for base in (SOURCE, STAGED):
        if os.path.exists(os.path.join(base, "code", "utils", "dycon_losses.py")):
            return base
    return None


def available() -> bool:
    return root() is not None


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def dycon_losses():
    """The reference's code/utils/dycon_losses.py, imported by path (it needs only math + torch)."""
    base = root()
    if base is None:
        raise FileNotFoundError("reference not found: neither /root/reference nor baseline/_ref is present")
    return _load(os.path.join(base, "code", "utils", "dycon_losses.py"), "_ref_dycon_losses")


def stock_losses():
    """The reference's code/utils/losses.py (dice_loss, softmax_mse_loss, ...: needs torch + numpy)."""
    base = root()
    if base is None:
        raise FileNotFoundError("reference not found")
    return _load(os.path.join(base, "code", "utils", "losses.py"), "_ref_losses")


def unet3d():
    """The reference's UNet3D class (code/networks/UNet3D_contrastive.py:207-316), imported WITHOUT running
    code/networks/__init__.py (which imports the absent ``monai``): the module's own imports are
    ``networks.utils`` / ``networks.networks_other`` / ``networks.assp`` -- provided through a synthetic
    ``networks`` package that points at the same directory."""
    base = root()
    if base is None:
        raise FileNotFoundError("reference not found")
    ndir = os.path.join(base, "code", "networks")
    if "networks" not in sys.modules or getattr(sys.modules["networks"], "__path__", None) != [ndir]:
        pkg = types.ModuleType("networks")
        pkg.__path__ = [ndir]
        sys.modules["networks"] = pkg
    return _load(os.path.join(ndir, "UNet3D_contrastive.py"), "networks.UNet3D_contrastive").UNet3D

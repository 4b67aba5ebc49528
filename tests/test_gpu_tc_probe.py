"""Stand-alone hardware check of the tcgen05 / TMA / swizzle encodings (tests/native/tc_probe.cu)."""
import os
import subprocess

import pytest
import torch

from conftest import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.mark.parametrize("kc", [1, 2, 4])
def test_umma_descriptors_against_host_gemm(kc):
    from dycon_paper_replication_b200.csrc.build import build_probe
    exe = build_probe()
    out = subprocess.run([exe, str(kc)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("OK") == 2 and "FAIL" not in out.stdout, out.stdout

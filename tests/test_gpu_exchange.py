"""dycon_exchange_sums (NVLink peer-memory all-reduce of the partial sums) on ONE GPU: a world of 1 whose only
peer is the rank itself.  The multi-rank protocol is covered by tests/test_sharded_gloo.py (host logic, CPU,
gloo) and by tools/multi_gpu_check.py (run under torchrun on a multi-GPU box)."""
import ctypes

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def _call(L, local, out, table, seq, kind=0, scale=0.0, lam=0.0, loss=None):
    from dycon_paper_replication_b200 import _lib
    _lib.check(L.dycon_exchange_sums(ctypes.c_void_p(local.data_ptr()), local.numel(), ctypes.c_void_p(out.data_ptr()),
                                     table, 0, 1, ctypes.c_void_p(seq.data_ptr()), kind, scale, lam,
                                     ctypes.c_void_p(loss.data_ptr()) if loss is not None else None, -1.0,
                                     torch.cuda.current_stream().cuda_stream), "dycon_exchange_sums")


def test_self_exchange_and_fused_losses():
    from dycon_paper_replication_b200 import _lib
    L = _lib.lib()
    inbox = torch.zeros(L.dycon_exchange_inbox_bytes() // 8, dtype=torch.float64, device="cuda")
    seq = torch.zeros(_lib.EXCHANGE_CHANNELS, dtype=torch.int64, device="cuda")
    table = (ctypes.c_void_p * 1)(inbox.data_ptr())
    assert L.dycon_exchange_enable_peer(torch.cuda.current_device()) == 0
    for it in range(9):                       # both slots, several rounds, every payload size
        n = 1 + it % 7
        x = torch.arange(1, n + 1, dtype=torch.float64, device="cuda") * (it + 1.5)
        out = torch.empty_like(x)
        _call(L, x, out, table, seq)
        assert torch.equal(out, x)
    assert int(seq[0].item()) == 9 and int(seq[1:].abs().sum().item()) == 0
    sums = torch.tensor([6.0, -3.0, 4.0], dtype=torch.float64, device="cuda")
    loss = torch.empty((), dtype=torch.float32, device="cuda")
    _call(L, sums[:1].clone(), torch.empty(1, dtype=torch.float64, device="cuda"), table, seq, _lib.EXCHANGE_UNCL, 0.25, 0.0, loss)
    assert loss.item() == pytest.approx(1.5)
    _call(L, sums.clone(), torch.empty(3, dtype=torch.float64, device="cuda"), table, seq, _lib.EXCHANGE_FECL, 0.5, 2.0, loss)
    assert loss.item() == pytest.approx(3.0)
    _call(L, sums.clone(), torch.empty(3, dtype=torch.float64, device="cuda"), table, seq, _lib.EXCHANGE_FECL_TEACHER, 0.5, 2.0, loss)
    assert loss.item() == pytest.approx(3.0 + 2.0 * (-3.0 / 4.0))


def test_rejects_bad_arguments():
    from dycon_paper_replication_b200 import _lib
    L = _lib.lib()
    x = torch.zeros(3, dtype=torch.float64, device="cuda")
    seq = torch.zeros(_lib.EXCHANGE_CHANNELS, dtype=torch.int64, device="cuda")
    table = (ctypes.c_void_p * 1)(x.data_ptr())
    stream = torch.cuda.current_stream().cuda_stream
    p = ctypes.c_void_p
    assert L.dycon_exchange_sums(p(x.data_ptr()), 8, p(x.data_ptr()), table, 0, 1, p(seq.data_ptr()), 0, 0.0, 0.0, None, -1.0, stream) < 0
    assert L.dycon_exchange_sums(p(x.data_ptr()), 3, p(x.data_ptr()), table, 1, 1, p(seq.data_ptr()), 0, 0.0, 0.0, None, -1.0, stream) < 0
    assert L.dycon_exchange_sums(p(x.data_ptr()), 1, p(x.data_ptr()), table, 0, 1, p(seq.data_ptr()), 3, 0.0, 0.0, None, -1.0, stream) < 0
    assert b"exchange" in L.dycon_last_error()

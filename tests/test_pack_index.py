"""CPU restatement of the index arithmetic of `pack16_kernel` (csrc/fecl_tc.cu): the swizzled 64 x 64 shared-memory tile,
both loaders (16-byte and scalar) and the transposed 16-byte writes.  It checks that the mapping is a plain transposition
with zero padding for even and odd shapes, that a float4 of four consecutive n stays one aligned 16-byte unit, and that the
transposed reads of a warp touch 32 distinct banks.  (The kernel itself is checked on the GPU against the oracle through every
FeCL test; this guards the swizzle on machines without one.)"""
import numpy as np
import pytest


def tix(d, n):
    return d * 64 + (n ^ (((d >> 3) & 7) << 2))


def emulate(N, D, vector):
    npad, dpad = (N + 127) // 128 * 128, (D + 63) // 64 * 64
    src = np.random.default_rng(0).standard_normal((D, N)).astype(np.float32)      # [d][n]: the caller's (D*N, 1, N) layout
    dst = np.full((npad, dpad), np.nan, np.float32)
    for n0 in range(0, npad, 64):
        for d0 in range(0, dpad, 64):
            tile = np.full(64 * 64, np.nan, np.float32)
            for tid in range(256):
                if vector:
                    for k in range(4):
                        idx = tid + 256 * k
                        d, n = d0 + (idx >> 4), n0 + (idx & 15) * 4
                        a = tix(idx >> 4, (idx & 15) * 4)
                        assert a % 4 == 0                      # one aligned 16-byte shared-memory store
                        tile[a:a + 4] = src[d, n:n + 4] if (n < N and d < D) else 0.0
                else:
                    tx, ty = tid & 63, tid >> 6
                    for k in range(ty, 64, 4):
                        n, d = n0 + tx, d0 + k
                        tile[tix(k, tx)] = src[d, n] if (n < N and d < D) else 0.0
            for tid in range(256):
                for k in range(2):
                    u = tid + 256 * k
                    g, r = u & 7, u >> 3
                    dst[n0 + r, d0 + 8 * g:d0 + 8 * g + 8] = [tile[tix(8 * g + j, r)] for j in range(8)]
    ref = np.zeros((npad, dpad), np.float32)
    ref[:N, :D] = src.T
    return dst, ref


@pytest.mark.parametrize("N,D,vector", [(128, 64, True), (80, 32, True), (203, 48, False), (257, 128, False), (100, 64, True)])
def test_pack_is_a_zero_padded_transposition(N, D, vector):
    dst, ref = emulate(N, D, vector)
    assert np.array_equal(dst, ref)


def test_transposed_reads_are_conflict_free():
    for warp in range(8):
        for k in range(2):
            for j in range(8):
                banks = set()
                for lane in range(32):
                    u = warp * 32 + lane + 256 * k
                    banks.add(tix(8 * (u & 7) + j, u >> 3) % 32)
                assert len(banks) == 32

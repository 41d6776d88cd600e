"""liblte_phy_rate_unmatch_turbo parity: CUDA gather-map kernel vs the oracle, bit-exact floats."""
import ctypes as C

import numpy as np
import pytest

from util import dptr, stream_ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [40, 48, 1024, 2240, 6144])
def test_rate_unmatch_matches_oracle(ctx, port, refo, K):
    import torch
    D = K + 4
    rng = np.random.default_rng(K)
    for rv in range(4):
        for E in (D, 2 * D + 3, 3 * D, 3 * D + 40, 6 * D + 8):
            for Ccb in (1, 2):
                n = 3
                e = (rng.standard_normal((n, E)) * 3).astype(np.float32)
                e[:, ::17] = 10000.0
                de = torch.from_numpy(e).cuda(); out = torch.empty((n, 3 * D), dtype=torch.float32, device="cuda")
                ctx.check(ctx.lib.rate_unmatch_turbo_batch(ctx.h, dptr(de), E, E, K, Ccb, 1, 250368, 8, 0, rv, n, dptr(out), stream_ptr()))
                got = out.cpu().numpy()
                for b in range(n):
                    want = port.rate_unmatch_turbo(e[b], D, rv=rv, C=Ccb)
                    assert np.array_equal(got[b], want), (K, rv, E, Ccb, b)
                if refo is not None and rv == 0:
                    phy = refo.RefPhy()
                    assert np.array_equal(got[0], phy.rate_unmatch_turbo(e[0], D, rv=rv, N_codeblocks=Ccb))


def test_round_trip_property(ctx, port):
    """rate_match -> rate_unmatch recovers every transmitted position's sign (the reference's manual test,
    liblte/manual_tests/ratematch_test.cc:59-251), here for every legal K at rv 0 and E = 3(K+4)."""
    import torch
    rng = np.random.default_rng(0)
    Ks = list(range(40, 512, 8)) + list(range(512, 1024, 16)) + list(range(1024, 2048, 32)) + list(range(2048, 6145, 64))
    for K in Ks[::5] + [6144]:
        D = K + 4; E = 3 * D
        d = rng.integers(0, 2, 3 * D).astype(np.uint8)
        e = port.rate_match_turbo(d, E, rv=0)
        llr = (1.0 - 2.0 * e.astype(np.float32))
        de = torch.from_numpy(llr).cuda(); out = torch.empty(3 * D, dtype=torch.float32, device="cuda")
        ctx.check(ctx.lib.rate_unmatch_turbo_batch(ctx.h, dptr(de), E, E, K, 1, 1, 250368, 8, 0, 0, 1, dptr(out), stream_ptr()))
        got = out.cpu().numpy().reshape(D, 3).T.reshape(-1)     # back to branch-major
        filled = got != 10000.0
        assert filled.all(), K
        assert np.array_equal((got < 0).astype(np.uint8), d), K


def test_host_twin(ctx, port):
    K = 1024; D = K + 4; E = 3 * D + 100
    rng = np.random.default_rng(9)
    e = rng.standard_normal(E).astype(np.float32)
    out = np.zeros(3 * D, np.float32); n = C.c_uint32(0)
    dummy = np.zeros(3 * D, np.uint8)
    rc = ctx.lib.rate_unmatch_turbo(ctx.h, e.ctypes.data_as(C.c_void_p), E, dummy.ctypes.data_as(C.c_void_p), D, 1, 1, 250368, 8, 0, 2,
                                    out.ctypes.data_as(C.c_void_p), C.byref(n))
    ctx.check(rc)
    assert n.value == 3 * D
    assert np.array_equal(out, port.rate_unmatch_turbo(e, D, rv=2))

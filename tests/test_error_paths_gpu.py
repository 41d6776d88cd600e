"""Argument validation and error behaviour of the C ABI on a live device: bad arguments must come back as LTEB200_ERR_ARGS /
LTEB200_ERR_UNSUPPORTED with a message, never as a crash or a silent success (the reference returns LIBLTE_ERROR_INVALID_INPUTS in
the same situations, liblte_phy.cc:3699-3705 and the entry checks of every liblte_phy_* function)."""
import ctypes as C

import numpy as np
import pytest

from util import dptr, stream_ptr
from openlte_b200.lib import Alloc

pytestmark = pytest.mark.gpu
ARGS, UNSUP = -1, -4


def test_bad_arguments_are_refused(ctx):
    import torch
    lib = ctx.lib
    z = torch.zeros(1 << 16, dtype=torch.float32, device="cuda")
    o8 = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    oi = torch.zeros(64, dtype=torch.int32, device="cuda")
    sn = np.array([3], np.int32); st = np.array([0], np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    # turbo: illegal block size, NULL pointers
    assert lib.turbo_decode_batch(ctx.h, dptr(z), 41, 1, 0, 0, dptr(o8), stream_ptr()) == ARGS
    assert lib.turbo_decode_batch(ctx.h, None, 40, 1, 0, 0, dptr(o8), stream_ptr()) == ARGS
    assert b"" != ctx.lib.last_error(ctx.h)
    # front end: subframe number out of range, NULL outputs
    assert lib.get_dl_subframe_and_ce_batch(ctx.h, dptr(z), 0, 30720, 4, 100, 0, 1, p(st), p(np.array([10], np.int32)), 1, dptr(z), dptr(z), stream_ptr()) == ARGS
    assert lib.get_dl_subframe_and_ce_batch(ctx.h, dptr(z), 0, 30720, 4, 100, 0, 1, p(st), p(sn), 1, None, dptr(z), stream_ptr()) == ARGS
    # PDSCH: allocation outside the band, transport block that needs more than one code block (the reference is broken there: refused)
    bad = Alloc.make(1, [98, 99, 100], 144, 1)
    assert lib.pdsch_decode_batch(ctx.h, dptr(z), dptr(z), 100, 0, 1, p(sn), 1, 2, C.byref(bad), 1, 0, 8, dptr(o8), 144, dptr(oi), stream_ptr()) == ARGS
    big = Alloc.make(3, list(range(50)), 20616, 1)
    assert lib.pdsch_decode_batch(ctx.h, dptr(z), dptr(z), 100, 0, 1, p(sn), 1, 2, C.byref(big), 1, 0, 8, dptr(o8), 20616, dptr(oi), stream_ptr()) in (ARGS, UNSUP)
    # control channels: more than one port for PDCCH, bad cell id for PBCH
    assert lib.pdcch_decode_batch(ctx.h, dptr(z), dptr(z), 100, 0, 2, C.c_float(1.0), p(sn), 1, dptr(o8), stream_ptr()) == UNSUP
    assert lib.bch_decode_batch(ctx.h, dptr(z), dptr(z), 100, 504, 1, dptr(o8), stream_ptr()) == ARGS
    # transmitter: transport block size that needs filler bits
    odd = Alloc.make(1, [0, 1, 2, 3], 150, 1)
    assert lib.dl_transmit_batch(ctx.h, dptr(o8), 256, C.byref(odd), 1, 4, 100, 0, 1, p(sn), 1, 2, dptr(z), None, stream_ptr()) in (ARGS, UNSUP)
    # the context still works afterwards
    tb = torch.zeros((1, 1, 144), dtype=torch.uint8, device="cuda")
    grid = torch.zeros((1, 14, 1200, 2), dtype=torch.float32, device="cuda")
    good = Alloc.make(1, [0, 1, 2, 3], 144, 0xFFFF)
    ctx.check(lib.dl_transmit_batch(ctx.h, dptr(tb), 144, C.byref(good), 1, 4, 100, 0, 1, p(sn), 1, 2, dptr(grid), None, stream_ptr()))
    torch.cuda.synchronize()
    assert float(grid.abs().sum()) > 0


def test_allocation_table_cache_is_bounded(ctx):
    """The per-allocation device tables (RE lists, scrambling sequences) are cached by content; a long scan meets thousands of distinct
    DCI-derived allocations.  More distinct allocations than the cache holds: old entries are dropped, results stay right for
    allocations met before and after the eviction."""
    import torch
    lib = ctx.lib
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    sn = np.array([3], np.int32)
    rng = np.random.default_rng(5)
    tbs = 144
    bits = torch.from_numpy(rng.integers(0, 2, (1, 1, tbs)).astype(np.uint8)).cuda()
    first = Alloc.make(1, [0, 1, 2, 3], tbs, 0x4000)
    grid = torch.zeros((1, 16, 72, 2), dtype=torch.float32, device="cuda")
    ctx.check(lib.dl_transmit_batch(ctx.h, dptr(bits), tbs, C.byref(first), 1, 0, 6, 7, 1, p(sn), 1, 2, dptr(grid), None, stream_ptr()))
    ce = torch.zeros((1, 1, 14, 72, 2), dtype=torch.float32, device="cuda"); ce[..., 0] = 1.0
    out = torch.zeros((1, 1, tbs), dtype=torch.uint8, device="cuda"); ok = torch.zeros((1, 1), dtype=torch.int32, device="cuda")

    def decode(al):
        ok.zero_()
        ctx.check(lib.pdsch_decode_batch(ctx.h, dptr(grid), dptr(ce), 6, 7, 1, p(sn), 1, 2, C.byref(al), 1, 0, 8, dptr(out), tbs, dptr(ok), stream_ptr()))
        torch.cuda.synchronize()
        return int(ok.item()), out.cpu().numpy().copy()

    ok0, tb0 = decode(first)
    assert ok0 == 1 and np.array_equal(tb0, bits.cpu().numpy())
    for i in range(1300):                                   # distinct RNTIs = distinct scrambling sequences = distinct cache entries
        decode(Alloc.make(1, [0, 1, 2, 3], tbs, 0x5000 + i))
    ok1, tb1 = decode(first)                                # evicted meanwhile and rebuilt
    assert ok1 == 1 and np.array_equal(tb1, tb0)

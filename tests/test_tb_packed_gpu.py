"""Packed transport-block output (lteb200_set_tb_format(LTEB200_TB_PACKED)): eight bits per byte, first bit in the MSB, equal to np.packbits of the
one-bit-per-byte output; through the device-resident call and through the host pipeline; refused where it cannot apply."""
import ctypes as C

import numpy as np
import pytest

from util import dptr, stream_ptr

pytestmark = pytest.mark.gpu


def test_packed_equals_packbits_of_bytes(ctx, lib):
    import torch
    from openlte_b200 import synth
    from openlte_b200.lib import Alloc
    cell, cfi, n_sf = 31, 2, 300
    allocs = [dict(mod=3, prb=list(range(25 * a, 25 * a + 25)), tbs=6120, rnti=0x100 + a) for a in range(2)] + [dict(mod=2, prb=list(range(60, 72)), tbs=2024, rnti=0x7)]
    tx = synth.make_subframes(20, allocs, N_id_cell=cell, seed=2, noise=0.3, first_subfr=1)
    seg = tx["samples"].reshape(-1)
    stream = np.concatenate([np.tile(seg, n_sf // 20), synth.crs_only_tail(100, cell, (1 + n_sf) % 10, 4)[:8192]])
    sc = 100.0 / max(np.abs(stream.real).max(), np.abs(stream.imag).max())
    iq = np.stack([np.rint(stream.real * sc), np.rint(stream.imag * sc)], -1).astype(np.int8)
    start = (np.arange(n_sf, dtype=np.int64) * 30720); sn = ((1 + np.arange(n_sf)) % 10).astype(np.int32)
    ca = (Alloc * 3)(*[Alloc.make(a["mod"], a["prb"], a["tbs"], a["rnti"]) for a in allocs])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    d_iq = torch.from_numpy(iq).cuda()
    tb = torch.zeros((n_sf, 3, 6120), dtype=torch.uint8, device="cuda"); ok = torch.zeros((n_sf, 3), dtype=torch.int32, device="cuda")
    ctx.check(lib.dl_receive_batch(ctx.h, dptr(d_iq), 1, iq.shape[0], 4, 100, cell, 1, p(start), p(sn), n_sf, cfi, ca, 3, 1, 8, dptr(tb), 6120, dptr(ok), stream_ptr()))
    torch.cuda.synchronize()
    want = tb.cpu().numpy(); okw = ok.cpu().numpy()
    assert okw.all()
    ctx.check(lib.set_tb_format(ctx.h, 1))
    try:
        tbp = torch.zeros((n_sf, 3, 768), dtype=torch.uint8, device="cuda"); ok2 = torch.zeros((n_sf, 3), dtype=torch.int32, device="cuda")
        ctx.check(lib.dl_receive_batch(ctx.h, dptr(d_iq), 1, iq.shape[0], 4, 100, cell, 1, p(start), p(sn), n_sf, cfi, ca, 3, 1, 8, dptr(tbp), 768, dptr(ok2), stream_ptr()))
        torch.cuda.synchronize()
        got = tbp.cpu().numpy()
        assert np.array_equal(ok2.cpu().numpy(), okw)
        for a, al in enumerate(allocs):
            nb = al["tbs"] // 8
            assert np.array_equal(got[:, a, :nb], np.packbits(want[:, a, : al["tbs"]], axis=-1)), f"allocation {a}"
        # host pipeline (chunked, pinned buffers)
        h_iq = torch.from_numpy(iq).pin_memory(); h_tb = torch.zeros((n_sf, 3, 768), dtype=torch.uint8).pin_memory(); h_ok = torch.zeros((n_sf, 3), dtype=torch.int32).pin_memory()
        ctx.check(lib.dl_receive_host(ctx.h, C.c_void_p(h_iq.data_ptr()), 1, iq.shape[0], 4, 100, cell, 1, p(start), p(sn), n_sf, cfi, ca, 3, 1, 8,
                                      C.c_void_p(h_tb.data_ptr()), 768, C.c_void_p(h_ok.data_ptr())))
        for a, al in enumerate(allocs):          # (bytes past a transport block's end are not written)
            assert np.array_equal(h_tb.numpy()[:, a, : al["tbs"] // 8], got[:, a, : al["tbs"] // 8]), f"host pipeline, allocation {a}"
        assert np.array_equal(h_ok.numpy(), okw)
        # the reference-exact decoder packs as well (same symbol-level kernel, its own decoder): equal to the bytes it produces unpacked
        lib.set_tb_format(ctx.h, 0)
        tba = torch.zeros((n_sf, 3, 6120), dtype=torch.uint8, device="cuda"); oka = torch.zeros((n_sf, 3), dtype=torch.int32, device="cuda")
        ctx.check(lib.dl_receive_batch(ctx.h, dptr(d_iq), 1, iq.shape[0], 4, 100, cell, 1, p(start), p(sn), n_sf, cfi, ca, 3, 0, 8, dptr(tba), 6120, dptr(oka), stream_ptr()))
        ctx.check(lib.set_tb_format(ctx.h, 1))
        ctx.check(lib.dl_receive_batch(ctx.h, dptr(d_iq), 1, iq.shape[0], 4, 100, cell, 1, p(start), p(sn), n_sf, cfi, ca, 3, 0, 8, dptr(tbp), 768, dptr(ok2), stream_ptr()))
        torch.cuda.synchronize()
        wa = tba.cpu().numpy(); ga = tbp.cpu().numpy()
        assert np.array_equal(ok2.cpu().numpy(), oka.cpu().numpy())
        for a, al in enumerate(allocs):
            assert np.array_equal(ga[:, a, : al["tbs"] // 8], np.packbits(wa[:, a, : al["tbs"]], axis=-1)), f"reference-exact decoder, allocation {a}"
        # refused: a block size that is not a multiple of 32 with the max-log-MAP decoder
        odd = (Alloc * 1)(Alloc.make(1, list(range(0, 10)), 464, 0x9))          # K = 488: not a multiple of 32
        assert lib.dl_receive_batch(ctx.h, dptr(d_iq), 1, iq.shape[0], 4, 100, cell, 1, p(start), p(sn), 4, cfi, odd, 1, 1, 8, dptr(tbp), 768, dptr(ok2), stream_ptr()) == -4
    finally:
        lib.set_tb_format(ctx.h, 0)

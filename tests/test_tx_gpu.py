"""Transmit chain on the GPU against the reference's own transmitter (oracle/_ref, where built), against the
numpy generator that tests/test_synth.py pins on the reference, and as a full-size round trip through the receive chain."""
import ctypes as C

import numpy as np
import pytest

from util import dptr, stream_ptr
from openlte_b200 import synth
from openlte_b200.lib import Alloc

pytestmark = pytest.mark.gpu


def gpu_tx(ctx, tb, allocs, cell, nrb, fs_idx, subfrs, cfi=2):
    """tb: [n_sf][n_alloc][stride] uint8 -> (grid [n_sf][14][nsc] complex64, samples [n_sf][n] complex64)"""
    import torch
    n_sf, n_alloc, stride = tb.shape
    d_tb = torch.from_numpy(np.ascontiguousarray(tb)).cuda()
    grid = torch.zeros((n_sf, 14, 12 * nrb, 2), dtype=torch.float32, device="cuda")
    iq = torch.zeros((n_sf, 1920 << fs_idx, 2), dtype=torch.float32, device="cuda")
    ca = (Alloc * n_alloc)(*[Alloc.make(a["mod"], a["prb"], a["tbs"], a["rnti"], rv_idx=a.get("rv", 0)) for a in allocs])
    sn = np.asarray(subfrs, np.int32)
    ctx.check(ctx.lib.dl_transmit_batch(ctx.h, dptr(d_tb), stride, ca, n_alloc, fs_idx, nrb, cell, 1, sn.ctypes.data_as(C.c_void_p), n_sf, cfi, dptr(grid),
                                        dptr(iq), stream_ptr()))
    torch.cuda.synchronize()
    g = grid.cpu().numpy(); s = iq.cpu().numpy()
    return (g[..., 0] + 1j * g[..., 1]).astype(np.complex64), (s[..., 0] + 1j * s[..., 1]).astype(np.complex64)


CASES = [(0, 100, 4, [dict(mod=3, prb=list(range(12)), tbs=2216, rnti=0x1234)], [1, 5, 0]),
         (77, 100, 4, [dict(mod=1, prb=list(range(40, 76)), tbs=3240, rnti=0x0A0B), dict(mod=2, prb=list(range(2, 20)), tbs=3112, rnti=0x7FFF, rv=2)], [5, 6]),
         (301, 25, 2, [dict(mod=1, prb=list(range(4, 10)), tbs=208, rnti=77), dict(mod=3, prb=list(range(12, 25)), tbs=1544, rnti=78)], [0, 3, 9]),
         (150, 6, 0, [dict(mod=2, prb=[1, 2, 3, 4], tbs=488, rnti=0xFFFF)], [2])]


@pytest.mark.parametrize("cell,nrb,fs_idx,allocs,subfrs", CASES)
def test_grid_and_samples_against_numpy_generator_and_reference(ctx, refo, cell, nrb, fs_idx, allocs, subfrs):
    rng = np.random.default_rng(cell)
    n_sf = len(subfrs)
    stride = max(a["tbs"] for a in allocs) + 8
    tb = np.zeros((n_sf, len(allocs), stride), np.uint8)
    for a, al in enumerate(allocs):
        tb[:, a, :al["tbs"]] = rng.integers(0, 2, (n_sf, al["tbs"]))
    g, s = gpu_tx(ctx, tb, allocs, cell, nrb, fs_idx, subfrs)
    want = synth.make_subframes(n_sf, allocs, N_id_cell=cell, N_rb_dl=nrb, fs_idx=fs_idx, cfi=2, subfr_nums=np.asarray(subfrs),
                                tb_bits=[tb[:, a, :al["tbs"]] for a, al in enumerate(allocs)])
    assert np.array_equal(g, want["grid"]), "resource grid differs from the generator that is pinned on the reference transmitter"
    assert np.abs(s - want["samples"]).max() <= 2e-6 * np.abs(want["samples"]).max()
    if refo is not None and fs_idx == 4:
        phy = refo.RefPhy(N_id_cell=cell, N_rb_dl=nrb)
        for b, sn in enumerate(subfrs):
            sf = refo.Subframe(); sf.clear(); sf.set_num(sn)
            phy.map_crs(sf, cell, 1)
            for a, al in enumerate(allocs):
                phy.pdsch_encode(sf, 2, cell, 1, al["mod"], al["prb"], al["tbs"], al["rnti"], tb[b, a, :al["tbs"]], rv=al.get("rv", 0))
            ref_grid = (sf.tx_symb_re[0, :14, :12 * nrb] + 1j * sf.tx_symb_im[0, :14, :12 * nrb]).astype(np.complex64)
            assert np.array_equal(g[b], ref_grid), "resource grid differs from the reference transmitter's"
            i_s, q_s = phy.create_dl_subframe(sf, 0, 30720)
            z = i_s + 1j * q_s
            assert np.abs(s[b] - z).max() <= 2e-6 * np.abs(z).max()


def test_full_size_round_trip(ctx):
    """BASELINE's full batch as a property test: 4000 distinct subframes x 4 transport blocks (64QAM, K = 6144) generated on the
    device, received by the fused chain: every CRC passes and every bit comes back (both decoders)."""
    import torch
    n_sf, cell, cfi = 4000, 17, 2
    allocs = [dict(mod=3, prb=list(range(25 * a, 25 * a + 25)), tbs=6120, rnti=0x100 + a) for a in range(4)]
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    tb = torch.randint(0, 2, (n_sf + 1, 4, 6120), dtype=torch.uint8, device="cuda", generator=g)
    iq = torch.zeros((n_sf + 1, 30720, 2), dtype=torch.float32, device="cuda")
    ca = (Alloc * 4)(*[Alloc.make(a["mod"], a["prb"], a["tbs"], a["rnti"]) for a in allocs])
    sn = ((1 + np.arange(n_sf + 1)) % 10).astype(np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx.check(ctx.lib.dl_transmit_batch(ctx.h, dptr(tb), 6120, ca, 4, 4, 100, cell, 1, p(sn), n_sf + 1, cfi, None, dptr(iq), stream_ptr()))
    start = (np.arange(n_sf) * 30720).astype(np.int64)
    for mode in (1, 0):
        out = torch.zeros((n_sf, 4, 6120), dtype=torch.uint8, device="cuda"); ok = torch.zeros((n_sf, 4), dtype=torch.int32, device="cuda")
        ctx.check(ctx.lib.dl_receive_batch(ctx.h, dptr(iq), 0, (n_sf + 1) * 30720, 4, 100, cell, 1, p(start), p(sn[:n_sf]), n_sf, cfi, ca, 4, mode, 8, dptr(out),
                                           6120, dptr(ok), stream_ptr()))
        torch.cuda.synchronize()
        assert bool(ok.all()) and bool((out == tb[:n_sf]).all()), mode

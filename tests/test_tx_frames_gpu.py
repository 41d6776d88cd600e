"""Device-side signal generator with everything LTE_fdd_dl_file_gen transmits (lteb200_dl_transmit_frames_batch): CRS of every port, PSS/SSS, PBCH,
PCFICH + PDCCH (DCI 1A), PDSCH with transmit diversity -- resource grids bit-identical to the REFERENCE's transmit functions
(tests/golden/tx_frames_ref.npz, made by tests/golden/make_golden_tx.py from liblte_phy_map_crs / map_pss / map_sss / bch_channel_encode /
pdcch_channel_encode / pdsch_channel_encode), and a closed loop: a generated stream goes through the scan path (PDCCH blind decode -> DCI ->
PDSCH) and comes back with the transmitted transport blocks."""
import ctypes as C
import os

import numpy as np
import pytest

from util import dptr, stream_ptr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tx_frames_ref.npz")
FS_IDX = {100: 4, 50: 3, 25: 2}


def run_case(ctx, lib, g, name, want_iq=False, gains=None):
    import torch
    from openlte_b200.lib import Alloc, TxControl
    n_rb, cell, n_ant, cfi, sfn0, pdcch = (int(x) for x in g[f"{name}_meta"])
    subframes = g[f"{name}_subframes"]; allocs = g[f"{name}_allocs"]; tb = g[f"{name}_tb"]; mib = np.ascontiguousarray(g[f"{name}_mib"])
    n_sf, n_alloc = subframes.shape[0], allocs.shape[0]
    # the generator numbers its frames by the subframe-0 crossings of the batch: feed it consecutive subframe numbers per frame as recorded
    sn = np.ascontiguousarray(subframes[:, 1], np.int32)
    ca = (Alloc * n_alloc)(*[Alloc.make(int(a[3]), list(range(int(a[4]), int(a[4] + a[5]))), int(a[2]), int(a[0]), tx_mode=1 if n_ant == 1 else 2) for a in allocs])
    ctl = TxControl()
    ctl.N_id_cell, ctl.N_ant, ctl.sync, ctl.pbch, ctl.sfn_first, ctl.pdcch, ctl.phich_res = cell, n_ant, 1, 1, sfn0, pdcch, 1.0
    for p in range(4):
        ctl.ant_gain_re[p] = 1.0 if gains is None else float(np.real(gains[p])); ctl.ant_gain_im[p] = 0.0 if gains is None else float(np.imag(gains[p]))
    ctl.mib_bits = mib.ctypes.data_as(C.c_void_p).value
    for a in range(n_alloc):
        ctl.mcs[a] = int(allocs[a][1])
    d_tb = torch.from_numpy(np.ascontiguousarray(tb)).cuda()
    nsc = 12 * n_rb
    d_grid = torch.zeros((n_sf, n_ant, 14, nsc, 2), dtype=torch.float32, device="cuda")
    n_s = 1920 << FS_IDX[n_rb]
    d_iq = torch.zeros((n_sf * n_s, 2), dtype=torch.float32, device="cuda") if want_iq else None
    n_symbs = cfi + (1 if n_rb <= 10 else 0)
    ctx.check(lib.dl_transmit_frames_batch(ctx.h, dptr(d_tb), tb.shape[2], ca, n_alloc, FS_IDX[n_rb], n_rb, C.byref(ctl), sn.ctypes.data_as(C.c_void_p), n_sf,
                                           n_symbs, dptr(d_grid), dptr(d_iq) if want_iq else None, stream_ptr()))
    torch.cuda.synchronize()
    grid = d_grid.cpu().numpy()
    return grid[..., 0] + 1j * grid[..., 1], (d_iq.cpu().numpy() if want_iq else None)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_grids_bit_identical_to_the_reference_transmitter(ctx, lib, name):
    g = np.load(GOLD)
    got, _ = run_case(ctx, lib, g, name)
    want = g[f"{name}_grid"]
    assert got.shape == want.shape
    bad = np.argwhere(got.astype(np.complex64) != want)
    assert bad.size == 0, f"case {name}: {bad.shape[0]} resource elements differ, first (subframe, port, symbol, subcarrier) = {bad[0]}: {got[tuple(bad[0])]} vs {want[tuple(bad[0])]}"


def test_generated_stream_through_the_scan_path(ctx, lib):
    """One port, 20 MHz: 40 consecutive subframes with PCFICH + 3 DCIs + PDSCH generated on the device, received by lteb200_dl_scan_batch
    (front end -> PDCCH blind decode -> DCI unpack -> PDSCH receive of every announced allocation): every transport block comes back."""
    import torch
    from openlte_b200.lib import Alloc, TxControl
    g = np.load(GOLD)
    n_rb, cell, cfi = 100, 17, 2
    allocs = g["a_allocs"]
    n_alloc, n_sf = allocs.shape[0], 40
    rng = np.random.default_rng(9)
    tb = np.zeros((n_sf + 1, n_alloc, 2216), np.uint8)
    for a in range(n_alloc):
        tb[:, a, : allocs[a][2]] = rng.integers(0, 2, (n_sf + 1, allocs[a][2]))
    sn = ((3 + np.arange(n_sf + 1)) % 10).astype(np.int32)
    ca = (Alloc * n_alloc)(*[Alloc.make(int(a[3]), list(range(int(a[4]), int(a[4] + a[5]))), int(a[2]), int(a[0])) for a in allocs])
    n_frames = 6
    mib = rng.integers(0, 2, (n_frames, 24)).astype(np.uint8)
    ctl = TxControl()
    ctl.N_id_cell, ctl.N_ant, ctl.sync, ctl.pbch, ctl.sfn_first, ctl.pdcch, ctl.phich_res = cell, 1, 1, 1, 0, 1, 1.0
    ctl.ant_gain_re[0] = 1.0
    ctl.mib_bits = mib.ctypes.data_as(C.c_void_p).value
    for a in range(n_alloc):
        ctl.mcs[a] = int(allocs[a][1])
    d_tb = torch.from_numpy(tb).cuda()
    d_iq = torch.zeros(((n_sf + 1) * 30720, 2), dtype=torch.float32, device="cuda")
    ctx.check(lib.dl_transmit_frames_batch(ctx.h, dptr(d_tb), 2216, ca, n_alloc, 4, n_rb, C.byref(ctl), sn.ctypes.data_as(C.c_void_p), n_sf + 1, cfi, None, dptr(d_iq),
                                           stream_ptr()))
    torch.cuda.synchronize()
    from openlte_b200.lib import ScanResult
    res = (ScanResult * n_sf)()
    cap = n_sf * 6 * 2216
    out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    used = C.c_longlong(0)
    start = (np.arange(n_sf, dtype=np.int64) * 30720)
    ctx.check(lib.dl_scan_batch(ctx.h, dptr(d_iq), 0, int(d_iq.shape[0]), 4, n_rb, cell, 1, C.c_float(1.0), start.ctypes.data_as(C.c_void_p),
                                sn[:n_sf].ctypes.data_as(C.c_void_p), n_sf, 1, 8, res, dptr(out), cap, C.byref(used), stream_ptr()))
    torch.cuda.synchronize()
    bits = out.cpu().numpy()
    for b in range(n_sf):
        r = res[b]
        # (a DCI whose neighbouring CCEs are empty is found again at aggregation level 8, exactly as the reference's blind decoder finds the
        # shipped capture's SIB1 twice: at least n_alloc entries, every allocation among them)
        assert r.cfi == cfi and r.N_alloc >= n_alloc, f"subframe {b}: cfi {r.cfi}, {r.N_alloc} allocations (pdcch_err {r.pdcch_err})"
        seen = set()
        for k in range(r.N_alloc):
            d = r.dci[k]
            a = [i for i in range(n_alloc) if int(allocs[i][0]) == d.alloc.rnti][0]
            seen.add(a)
            assert d.alloc.tbs == int(allocs[a][2]) and r.crc_ok[k] == 1
            assert np.array_equal(bits[r.tb_offset[k]: r.tb_offset[k] + d.alloc.tbs], tb[b, a, : d.alloc.tbs])
        assert len(seen) == n_alloc

"""lteb200_dl_scan_batch: the reference's scan loop (front end -> PDCCH -> PDSCH of every announced allocation) for a batch of
subframes in one call, against what the reference finds subframe by subframe (tests/golden/make_golden_control.py scan)."""
import ctypes as C

import numpy as np
import pytest

from util import dptr, stream_ptr
from openlte_b200.lib import ScanResult
from test_control_oracle import gold, alloc_row
from test_control_gpu import ERR_MAP

pytestmark = pytest.mark.gpu


def run_scan(ctx, iq_i8, fs_idx, nrb, cell, starts, subfrs, mode=0):
    import torch
    d_iq = torch.from_numpy(np.ascontiguousarray(iq_i8)).cuda()
    n = len(subfrs)
    res = (ScanResult * n)()
    tb = np.zeros(1 << 20, np.uint8); used = C.c_longlong(0)
    st = np.asarray(starts, np.int64); sn = np.asarray(subfrs, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx.check(ctx.lib.dl_scan_batch(ctx.h, dptr(d_iq), 1, iq_i8.shape[0], fs_idx, nrb, cell, 1, C.c_float(1.0), p(st), p(sn), n, mode, 8, res, p(tb), tb.size,
                                    C.byref(used), stream_ptr()))
    return res, tb, used.value


def check_against(g, tag, res, tb):
    for k in range(int(g[f"{tag}_n"])):
        err, cfi, nsym, nal = [int(x) for x in g[f"{tag}_hdr_{k}"]]
        R = res[k]
        assert (R.pdcch_err, R.cfi, R.N_symbs, R.N_alloc) == (ERR_MAP[err], cfi, nsym, nal), (tag, k)
        for a in range(nal):
            assert np.array_equal(alloc_row(R.dci[a].as_dict()), g[f"{tag}_allocs_{k}"][a]), (tag, k, a)
            ok = int(g[f"{tag}_ok_{k}"][a]) if err == 0 else 0
            assert R.crc_ok[a] == ok, (tag, k, a)
            if ok:
                want = g[f"{tag}_bits_{k}_{a}"]
                assert R.tb_offset[a] >= 0 and np.array_equal(tb[R.tb_offset[a]:R.tb_offset[a] + want.size], want), (tag, k, a)
            else:
                assert R.tb_offset[a] == -1


@pytest.mark.parametrize("mode", [0, 1])
def test_scan_of_the_shipped_capture(ctx, mode):
    """BASELINE config 1: subframes 1-9 of the first frame of the reference's capture in one call: only subframe 5 carries control
    information -- SIB1's DCI 1A, found twice (aggregation level 4 and 8), decoded twice, like the reference."""
    g = gold("control_scan_ref.npz"); cap = gold("capture_cfg1.npz")
    fs0 = int(g["cap_frame_start"])
    sn = list(range(1, 10))
    res, tb, used = run_scan(ctx, cap["iq"], 4, 100, 0, [fs0 + s * 30720 for s in sn], sn, mode)
    check_against(g, "cap", res, tb)
    assert used == 2 * 144 and np.array_equal(tb[:144], cap["sib1_bits"])


def test_scan_of_a_synthetic_stream_from_the_reference_transmitter(ctx):
    g = gold("control_scan_ref.npz")
    nrb, cell, fs_idx = [int(x) for x in g["syn_cfg"]]
    n_s = 1920 << fs_idx
    sn = [1, 2, 3, 4]
    res, tb, used = run_scan(ctx, g["syn_iq"], fs_idx, nrb, cell, [s * n_s for s in sn], sn)
    check_against(g, "syn", res, tb)
    # and the same subframes in a different order, three times over: grouping by allocation must not depend on position
    order = [3, 1, 4, 2, 2, 4, 1, 3, 1, 2, 3, 4]
    res2, tb2, used2 = run_scan(ctx, g["syn_iq"], fs_idx, nrb, cell, [s * n_s for s in order], order)
    assert used2 == 3 * used
    for pos, s in enumerate(order):
        A, B = res2[pos], res[s - 1]
        assert (A.pdcch_err, A.cfi, A.N_alloc, list(A.crc_ok)) == (B.pdcch_err, B.cfi, B.N_alloc, list(B.crc_ok))
        for a in range(A.N_alloc):
            if A.crc_ok[a]:
                n = A.dci[a].alloc.tbs
                assert np.array_equal(tb2[A.tb_offset[a]:A.tb_offset[a] + n], tb[B.tb_offset[a]:B.tb_offset[a] + n])

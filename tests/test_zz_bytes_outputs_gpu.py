"""Entry points whose outputs are DEFINED as one bit per byte keep that format when the context is set to packed transport blocks
(lteb200_set_tb_format): the host twin of liblte_phy_pdsch_channel_decode (liblte_phy.h:906-913), here on BASELINE config 1's SIB1.
(Written after this round's GPU budget was spent, so it sorts last: everything before it ran on a B200 with the final kernels.)"""
import ctypes as C
import os

import numpy as np
import pytest

from openlte_b200.lib import Alloc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
p = lambda a: a.ctypes.data_as(C.c_void_p)


def test_host_twin_returns_bytes_under_a_packed_context(ctx):
    z = np.load(os.path.join(GOLD, "capture_cfg1.npz"))
    cap = {k: z[k] for k in z.files}
    i_s = np.ascontiguousarray(cap["iq"][:, 0].astype(np.float32)); q_s = np.ascontiguousarray(cap["iq"][:, 1].astype(np.float32))
    cell = 3 * int(cap["N_id_1"]) + int(cap["N_id_2"])
    rx_re = np.zeros((16, 1200), np.float32); rx_im = np.zeros((16, 1200), np.float32)
    ce_re = np.zeros((4, 16, 1200), np.float32); ce_im = np.zeros((4, 16, 1200), np.float32)
    ctx.check(ctx.lib.get_dl_subframe_and_ce(ctx.h, p(i_s), p(q_s), i_s.size, 4, 100, int(cap["sib1_frame_start"]), 5, cell, 1, p(rx_re), p(rx_im), p(ce_re), p(ce_im)))
    al = Alloc.make(int(cap["mod"]), [int(x) for x in cap["prb"]], int(cap["tbs"]), int(cap["rnti"]), rv_idx=int(cap["rv"]))
    ctx.check(ctx.lib.set_tb_format(ctx.h, 1))
    try:
        for mode in (0, 1):
            out = np.full(int(cap["tbs"]), 7, np.uint8); n = C.c_uint32(0)
            rc = ctx.lib.pdsch_channel_decode(ctx.h, p(rx_re), p(rx_im), p(ce_re), p(ce_im), 100, 5, C.byref(al), int(cap["cfi"]), cell, 1, mode, 8, p(out), C.byref(n))
            assert rc == 0 and n.value == int(cap["tbs"])
            assert np.array_equal(out, cap["sib1_bits"]), mode
    finally:
        ctx.check(ctx.lib.set_tb_format(ctx.h, 0))

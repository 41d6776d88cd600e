"""PBCH and PCFICH/PDCCH receive on the GPU through the C ABI: bit-exact against the CPU restatement
and against the fixtures the reference produced (tests/golden/make_golden_control.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from util import dptr, stream_ptr
from openlte_b200.lib import BchResult, PdcchRaw, Dci
from test_control_oracle import gold, planes, alloc_row

pytestmark = pytest.mark.gpu


def to_dev_grid(re, im):
    """[.., rows, nsc] float planes -> device float2 tensor of the same leading shape"""
    import torch
    return torch.from_numpy(np.ascontiguousarray(np.stack([re, im], -1).astype(np.float32))).cuda()


def test_bch_batch_and_host_twin(ctx, port):
    import torch
    g = gold("control_bch_ref.npz")
    for n in range(int(g["n_cases"])):
        nrb, cell, nant, sfn, err, na, off = [int(x) for x in g[f"cfg_{n}"]]
        # batch API: three copies of the subframe (the third with the grid negated -> different decode, same code path)
        gr = to_dev_grid(g[f"rx_re_{n}"], g[f"rx_im_{n}"])                       # [16][nsc][2]
        ce = to_dev_grid(g[f"ce_re_{n}"][:, :14], g[f"ce_im_{n}"][:, :14])       # [4][14][nsc][2]
        gr3 = torch.stack([gr, gr, -gr]).contiguous(); ce3 = torch.stack([ce, ce, ce]).contiguous()
        out = torch.zeros(3 * C.sizeof(BchResult), dtype=torch.uint8, device="cuda")
        ctx.check(ctx.lib.bch_decode_batch(ctx.h, dptr(gr3), dptr(ce3), nrb, cell, 3, dptr(out), stream_ptr()))
        torch.cuda.synchronize()
        res = (BchResult * 3).from_buffer_copy(out.cpu().numpy().tobytes())
        for b in range(2):
            assert res[b].ok == (1 if err == 0 else 0), (n, b)
            if err == 0:
                assert (res[b].N_ant, res[b].offset) == (na, off) and np.array_equal(np.frombuffer(bytes(res[b].bits), np.uint8), g[f"rx_bits_{n}"])
        e3, a3, b3, o3 = port.bch_channel_decode(planes(-g[f"rx_re_{n}"]), planes(-g[f"rx_im_{n}"]), planes(g[f"ce_re_{n}"]), planes(g[f"ce_im_{n}"]), nrb, cell)
        assert res[2].ok == (1 if e3 == 0 else 0)
        if e3 == 0:
            assert (res[2].N_ant, res[2].offset) == (a3, o3) and np.array_equal(np.frombuffer(bytes(res[2].bits), np.uint8), b3)
        # host twin, planes as the reference's subframe struct
        rr, ri, cr, ci = planes(g[f"rx_re_{n}"]), planes(g[f"rx_im_{n}"]), planes(g[f"ce_re_{n}"]), planes(g[f"ce_im_{n}"])
        N_ant = C.c_uint8(9); offset = C.c_uint8(9); nb = C.c_uint32(0); bits = np.zeros(24, np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = ctx.lib.bch_channel_decode(ctx.h, p(rr), p(ri), p(cr), p(ci), nrb, cell, C.byref(N_ant), p(bits), C.byref(nb), C.byref(offset))
        if err == 0:
            assert rc == 0 and (N_ant.value, offset.value, nb.value) == (na, off, 24) and np.array_equal(bits, g[f"tx_bits_{n}"])
        else:
            assert rc == -3 and N_ant.value == 0


def gpu_pdcch(ctx, rx_re, rx_im, ce_re, ce_im, nrb, cell, subfrs):
    """rx_*: [B][16][nsc], ce_*: [B][14][nsc] -> list of PdcchRaw"""
    import torch
    B = len(subfrs)
    gr = to_dev_grid(rx_re, rx_im); ce = to_dev_grid(ce_re, ce_im)
    out = torch.zeros(B * C.sizeof(PdcchRaw), dtype=torch.uint8, device="cuda")
    sn = np.asarray(subfrs, np.int32)
    ctx.check(ctx.lib.pdcch_decode_batch(ctx.h, dptr(gr), dptr(ce), nrb, cell, 1, C.c_float(1.0), sn.ctypes.data_as(C.c_void_p), B, dptr(out), stream_ptr()))
    torch.cuda.synchronize()
    return (PdcchRaw * B).from_buffer_copy(out.cpu().numpy().tobytes())


def unpack(ctx, raw, nrb):
    allocs = (Dci * 6)(); n = C.c_int(0)
    rc = ctx.lib.pdcch_unpack(C.byref(raw), nrb, 1, allocs, C.byref(n))
    return rc, [allocs[i].as_dict() for i in range(n.value)]


ERR_MAP = {0: 0, 1: -1, 3: -3, 4: -5}     # LIBLTE_ERROR_ENUM -> LTEB200_ERR_*


def test_pdcch_batch_against_oracle_and_reference(ctx, port):
    g = gold("control_pdcch_ref.npz")
    for n in range(int(g["n_cases"])):
        nrb, cell, cfi, subfr, err, rcfi, nsym, nal = [int(x) for x in g[f"cfg_{n}"]]
        rx_re, rx_im, ce_re, ce_im = g[f"rx_re_{n}"], g[f"rx_im_{n}"], g[f"ce_re_{n}"], g[f"ce_im_{n}"]
        # batch of two: the fixture, and the same grid decoded as if it were another subframe number (wrong scrambling)
        other = (subfr + 3) % 10
        raws = gpu_pdcch(ctx, np.stack([rx_re, rx_re]), np.stack([rx_im, rx_im]), np.stack([ce_re[:14], ce_re[:14]]), np.stack([ce_im[:14], ce_im[:14]]),
                         nrb, cell, [subfr, other])
        for raw, sn in zip(raws, (subfr, other)):
            r = port.pdcch_channel_decode(planes(rx_re), planes(rx_im), planes(ce_re), planes(ce_im), nrb, cell, 1, sn)
            assert (raw.cfi, raw.N_symbs) == (r["cfi"], r["N_symbs"]), (n, sn)
            if raw.cfi:
                assert list(raw.found) == r["raw_found"].tolist(), (n, sn, list(raw.found), r["raw_found"].tolist())
                assert list(raw.rnti) == r["raw_rnti"].tolist()
                for s in range(12):
                    if raw.found[s]:
                        assert np.array_equal(np.frombuffer(bytes(raw.bits[s]), np.uint8), r["raw_bits"][s]), (n, sn, s)
            rc, allocs = unpack(ctx, raw, nrb)
            assert rc == ERR_MAP[r["err"]], (n, sn, rc, r["err"])
            assert allocs == r["allocs"]
        # the reference's own result for the fixture subframe
        rc, allocs = unpack(ctx, raws[0], nrb)
        assert (rc, raws[0].cfi, raws[0].N_symbs, len(allocs)) == (ERR_MAP[err], rcfi, nsym, nal)
        for a, want in zip(allocs, g[f"allocs_{n}"]):
            assert np.array_equal(alloc_row(a), want)


def test_pdcch_host_twin_on_the_shipped_capture(ctx):
    """BASELINE config 1 through the host twin of liblte_phy_pdcch_channel_decode: SIB1's DCI, then PDSCH with that allocation."""
    g = gold("capture_cfg1.npz")
    ce_re = np.zeros((16, 1200), np.float32); ce_im = np.zeros((16, 1200), np.float32)
    ce_re[:14] = g["sib1_ce_re"]; ce_im[:14] = g["sib1_ce_im"]
    rr = np.ascontiguousarray(g["sib1_rx_re"], np.float32); ri = np.ascontiguousarray(g["sib1_rx_im"], np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cfi = C.c_uint32(0); ns = C.c_uint32(0); na = C.c_uint32(0); allocs = (Dci * 6)()
    rc = ctx.lib.pdcch_channel_decode(ctx.h, p(rr), p(ri), p(ce_re), p(ce_im), 100, 5, 0, 1, C.c_float(1.0), 0, C.byref(cfi), C.byref(ns), allocs, C.byref(na))
    assert rc == 0 and cfi.value == int(g["cfi"]) and na.value >= 1
    a = allocs[0].as_dict()
    assert (a["rnti"], a["tbs"], a["mod_type"], a["rv_idx"], a["prb"]) == (int(g["rnti"]), int(g["tbs"]), int(g["mod"]), int(g["rv"]), g["prb"].tolist())
    out = np.zeros(a["tbs"], np.uint8); nb = C.c_uint32(0)
    ctx.check(ctx.lib.pdsch_channel_decode(ctx.h, p(rr), p(ri), p(ce_re), p(ce_im), 100, 5, C.byref(allocs[0].alloc), ns.value, 0, 1, 0, 8, p(out), C.byref(nb)))
    assert nb.value == 144 and np.array_equal(out, g["sib1_bits"])
    # more than one antenna port is refused loudly (the reference is undefined there)
    assert ctx.lib.pdcch_channel_decode(ctx.h, p(rr), p(ri), p(ce_re), p(ce_im), 100, 5, 0, 2, C.c_float(1.0), 0, C.byref(cfi), C.byref(ns), allocs, C.byref(na)) == -4


def test_pdcch_large_batch_is_consistent(ctx):
    """4000 subframes (the bench batch) of the same control region under different subframe numbers: every subframe with the
    right number decodes the same DCI, the others find the same nothing as their single-subframe decode."""
    g = gold("control_pdcch_ref.npz")
    nrb, cell, cfi, subfr = [int(x) for x in g["cfg_1"][:4]]
    B = 4000
    rx_re = np.broadcast_to(g["rx_re_1"], (B, 16, 12 * nrb)); rx_im = np.broadcast_to(g["rx_im_1"], (B, 16, 12 * nrb))
    ce_re = np.broadcast_to(g["ce_re_1"][:14], (B, 14, 12 * nrb)); ce_im = np.broadcast_to(g["ce_im_1"][:14], (B, 14, 12 * nrb))
    sn = [(subfr + (b % 3)) % 10 for b in range(B)]
    raws = gpu_pdcch(ctx, rx_re, rx_im, ce_re, ce_im, nrb, cell, sn)
    base = {s: bytes(raws[i]) for i, s in list(enumerate(sn))[:3]}
    assert all(bytes(raws[b]) == base[sn[b]] for b in range(B))
    assert sum(raws[0].found) >= 3 and raws[0].cfi == cfi


def test_mib_of_the_shipped_capture(ctx):
    """BASELINE config 1: MIB from subframe 0 of the frame the cell search finds in the reference's own capture, through the batched
    path (int8 I/Q -> OFDM -> four-port estimate at N_rb_dl = 6 -> PBCH): bits 1010100..., one antenna port, SFN mod 4 = 1."""
    import torch
    g = gold("capture_cfg1.npz")
    d_iq = torch.from_numpy(np.ascontiguousarray(g["iq"])).cuda()
    start = np.array([int(g["frame_start_idx"])], np.int64); sn = np.zeros(1, np.int32)
    grid = torch.zeros((1, 16, 72, 2), dtype=torch.float32, device="cuda"); ce = torch.zeros((1, 4, 14, 72, 2), dtype=torch.float32, device="cuda")
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx.check(ctx.lib.get_dl_subframe_and_ce_batch(ctx.h, dptr(d_iq), 1, g["iq"].shape[0], 4, 6, 0, 4, p(start), p(sn), 1, dptr(grid), dptr(ce), stream_ptr()))
    out = torch.zeros(C.sizeof(BchResult), dtype=torch.uint8, device="cuda")
    ctx.check(ctx.lib.bch_decode_batch(ctx.h, dptr(grid), dptr(ce), 6, 0, 1, dptr(out), stream_ptr()))
    torch.cuda.synchronize()
    r = BchResult.from_buffer_copy(out.cpu().numpy().tobytes())
    assert (r.ok, r.N_ant, r.offset) == (1, int(g["mib_N_ant"]), int(g["mib_offset"]))
    assert np.array_equal(np.frombuffer(bytes(r.bits), np.uint8), g["mib_bits"])

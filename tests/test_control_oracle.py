"""Control channels: the CPU restatement (oracle/port/control.cc) against fixtures produced by the
reference itself (tests/golden/make_golden_control.py)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def planes(a, rows=16):
    """[.., rows, nsc] -> [.., 16, 1200] zero padded, the LIBLTE_PHY_SUBFRAME_STRUCT layout the oracle takes"""
    out = np.zeros(a.shape[:-2] + (16, 1200), np.float32)
    out[..., :a.shape[-2], :a.shape[-1]] = a
    return out


def alloc_row(a):
    r = np.zeros(228, np.int32)
    r[:8] = [a["rnti"], a["tbs"], a["mcs"], a["N_prb"], a["rv_idx"], a["mod_type"], a["tx_mode"], a["chan_type"]]
    r[8:8 + len(a["prb"])] = a["prb"]; r[118:118 + len(a["prb1"])] = a["prb1"]
    return r


def test_conv_code_pieces_and_dci_unpack(port):
    g = gold("control_pieces_ref.npz")
    assert [[port.tbs_1a(i, 2), port.tbs_1a(i, 3)] for i in range(27)] == g["tbs_1a"].tolist()
    assert [port.tbs_1c(i) for i in range(32)] == g["tbs_1c"].tolist()
    for n in range(int(g["n_cc"])):
        d = port.rate_unmatch_conv(g[f"cc_e_{n}"], int(g[f"cc_nc_{n}"]))
        assert np.array_equal(d, g[f"cc_d_{n}"]), n
        assert np.array_equal(port.viterbi_decode(d), g[f"cc_c_{n}"]), n      # including the reference's surplus trailing bits
    for row in g["dci_rows"]:
        N, fmt, rnti, err = [int(x) for x in row[:4]]
        bits = row[4:36].astype(np.uint8)
        e, a = (port.dci_1a_unpack if fmt == 0 else port.dci_1c_unpack)(bits, rnti, N, 1)
        assert (e != 0) == (err != 0), (N, fmt, e, err)
        if err == 0:
            assert np.array_equal(alloc_row(a), row[36:]), (N, fmt)


def test_bch_decode_against_reference(port):
    g = gold("control_bch_ref.npz")
    for n in range(int(g["n_cases"])):
        nrb, cell, nant, sfn, err, na, off = [int(x) for x in g[f"cfg_{n}"]]
        e, a, bits, o = port.bch_channel_decode(planes(g[f"rx_re_{n}"]), planes(g[f"rx_im_{n}"]), planes(g[f"ce_re_{n}"]), planes(g[f"ce_im_{n}"]), nrb, cell)
        assert (e, a) == (err, na), (n, e, a, err, na)
        if err == 0:
            assert o == off and np.array_equal(bits, g[f"rx_bits_{n}"]) and np.array_equal(bits, g[f"tx_bits_{n}"]) and a == nant and o == sfn % 4


def test_pdcch_decode_against_reference(port):
    g = gold("control_pdcch_ref.npz")
    for n in range(int(g["n_cases"])):
        nrb, cell, cfi, subfr, err, rcfi, nsym, nal = [int(x) for x in g[f"cfg_{n}"]]
        r = port.pdcch_channel_decode(planes(g[f"rx_re_{n}"]), planes(g[f"rx_im_{n}"]), planes(g[f"ce_re_{n}"]), planes(g[f"ce_im_{n}"]), nrb, cell, 1, subfr)
        assert (r["err"], r["cfi"], r["N_symbs"], len(r["allocs"])) == (err, rcfi, nsym, nal), (n, r["err"], r["cfi"], r["N_symbs"], len(r["allocs"]))
        for a, want in zip(r["allocs"], g[f"allocs_{n}"]):
            assert np.array_equal(alloc_row(a), want), n


def test_pdcch_of_the_shipped_capture(port):
    """BASELINE config 1: subframe 5 of the reference's own capture carries SIB1's DCI 1A (SI-RNTI, 4 PRB, TBS 144)."""
    g = gold("capture_cfg1.npz")
    ce_re = np.zeros((16, 1200), np.float32); ce_im = np.zeros((16, 1200), np.float32)
    ce_re[:14] = g["sib1_ce_re"]; ce_im[:14] = g["sib1_ce_im"]
    r = port.pdcch_channel_decode(g["sib1_rx_re"], g["sib1_rx_im"], ce_re, ce_im, 100, 0, 1, 5)
    assert r["err"] == 0 and r["cfi"] == int(g["cfi"]) and len(r["allocs"]) >= 1
    a = r["allocs"][0]
    assert (a["rnti"], a["tbs"], a["mod_type"], a["rv_idx"], a["prb"]) == (int(g["rnti"]), int(g["tbs"]), int(g["mod"]), int(g["rv"]), g["prb"].tolist())

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


def test_product_cce_map_equals_oracle_for_every_bandwidth(port):
    """The table the PDCCH kernel gathers from (host code inside libopenlte_b200.so, callable without a device) against the oracle's
    restatement of the reference's REG walk / shift / de-interleaver, for all six bandwidths, 1-4 control symbols, three PHICH
    resources and a spread of cell identities; plus the structural property: no resource element is used twice and none is a CRS,
    PCFICH or PHICH position."""
    import ctypes as C
    import openlte_b200
    lib = openlte_b200.load()
    out = np.zeros(576, np.int32)
    n_checked = 0
    for nrb in (6, 15, 25, 50, 75, 100):
        nsc = 12 * nrb
        for cell in (0, 1, 2, 5, 77, 150, 301, 503):
            for pres in (1 / 6, 1.0, 2.0):
                for ns in (1, 2, 3, 4):
                    want, n_reg = port.pdcch_cce_map(nrb, cell, 1, ns, pres)
                    assert lib.pdcch_cce_map(nrb, cell, 1, ns, C.c_float(pres), out.ctypes.data_as(C.c_void_p)) == 0
                    n_cce = want.shape[0]
                    got = out.reshape(16, 36)
                    for c in range(16):
                        if c < n_cce:
                            w = (want[c] // 1200) * nsc + (want[c] % 1200)          # the oracle addresses rows of 1200 subcarriers
                            assert np.array_equal(got[c], w), (nrb, cell, pres, ns, c)
                        else:
                            assert (got[c] == -1).all()
                    used = got[:min(n_cce, 16)].reshape(-1)
                    assert used.size == np.unique(used).size
                    sym, k = used // nsc, used % nsc
                    assert (sym < ns).all()
                    assert not ((sym == 0) & (k % 3 == cell % 3)).any()              # CRS positions of symbol 0 (both ports' offsets)
                    n_checked += 1
    assert n_checked == 6 * 8 * 3 * 4


def test_product_dci_unpack_against_reference_fixtures():
    """lteb200_pdcch_unpack is host code: the reference's dci_1a_unpack / dci_1c_unpack results on random payloads (fixtures) must come
    back from it field for field, without a device."""
    import ctypes as C
    import openlte_b200
    from openlte_b200.lib import PdcchRaw, Dci
    lib = openlte_b200.load()
    g = gold("control_pieces_ref.npz")
    n_ok = 0
    for row in g["dci_rows"]:
        N, fmt, rnti, err = [int(x) for x in row[:4]]
        raw = PdcchRaw(); raw.cfi = 2; raw.N_symbs = 2
        slot = fmt                                   # candidate 0, format 1A -> slot 0, 1C -> slot 1
        raw.found[slot] = 1; raw.rnti[slot] = rnti
        for i in range(32):
            raw.bits[slot][i] = int(row[4 + i])
        allocs = (Dci * 6)(); n = C.c_int(0)
        rc = lib.pdcch_unpack(C.byref(raw), N, 1, allocs, C.byref(n))
        assert (rc != 0) == (err != 0), (N, fmt, rc, err)
        if err == 0:
            assert n.value == 1 and np.array_equal(alloc_row(allocs[0].as_dict()), row[36:]), (N, fmt)
            n_ok += 1
    assert n_ok > 100


def test_control_channels_against_live_reference_random_configs(port, refo):
    """Randomised differential test against the reference itself (fresh draws, not the committed fixtures): PBCH with 1/2/4 ports at every
    bandwidth (liblte_phy.cc:3968-4100) and PCFICH + PDCCH blind decode with one to three DCI 1A messages (:4519-5134), the signal built by
    the reference's own transmit functions, grid and estimate by its front end.  Error code, port count, MIB bits, SFN offset, CFI,
    symbol count and every decoded allocation field of the port equal the reference's -- also where the reference fails."""
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    import sys
    from util import ref_loopback
    FS = {100: 30.72, 75: 30.72, 50: 15.36, 25: 7.68, 15: 3.84, 6: 1.92}
    rng = np.random.default_rng(4242)
    bands = (100, 75, 50, 25, 15, 6)
    n_mib = 0
    for t in range(30):
        nrb = bands[t % 6]; cell = int(rng.integers(0, 504)); nant = (1, 2, 4)[(t // 6 + t) % 3]; sfn = int(rng.integers(0, 1024))
        noise = float(rng.choice([0.0, 0.02, 0.05, 300.0], p=[0.3, 0.3, 0.3, 0.1]))
        bits = rng.integers(0, 2, 24).astype(np.uint8)
        L = ref_loopback(refo, 0, [], N_id_cell=cell, N_rb_dl=nrb, fs=FS[nrb], N_ant=nant, bch=dict(bits=bits, sfn=sfn), noise=noise, seed=1000 + t)
        sf = refo.Subframe(); sf.clear()
        L["phy"].get_dl_subframe_and_ce(L["i"], L["q"], 0, 0, cell, 4, sf)
        err, na, ob, off = L["phy"].bch_decode(sf, cell)
        e, a, got, o = port.bch_channel_decode(sf.rx_symb_re.copy(), sf.rx_symb_im.copy(), sf.rx_ce_re.copy(), sf.rx_ce_im.copy(), nrb, cell)
        assert (e, a) == (err, na), (t, nrb, cell, nant, noise)
        if err == 0:
            assert o == off and np.array_equal(got, np.asarray(ob[:24], np.uint8)), (t, nrb, cell, nant)
            n_mib += 1
    assert n_mib >= 15
    n_dci = 0
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)     # the reference prints a line per "format 0" DCI
    try:
        for t in range(48):
            nrb = bands[t % 6]; cell = int(rng.integers(0, 504)); subfr = int(rng.integers(0, 10))
            cfi = int(rng.integers(1, 4)) if nrb > 10 else int(rng.integers(2, 4))
            n_msg = int(rng.integers(0, 4)) if nrb >= 25 else int(rng.integers(0, 2))
            allocs, p0 = [], 0
            for m in range(n_msg):
                n_prb = int(rng.integers(1, max(2, nrb // 6)))
                allocs.append(dict(rnti=int(rng.choice([0xFFFF, 0xFFFE, int(rng.integers(1, 60))])), mcs=int(rng.integers(0, 10)), N_prb=n_prb, prb0=p0))
                p0 += n_prb
            noise = float(rng.choice([0.0, 0.03, 0.1, 0.6]))
            L = ref_loopback(refo, subfr, [], N_id_cell=cell, N_rb_dl=nrb, fs=FS[nrb], N_ant=1, pdcch=dict(cfi=cfi, allocs=allocs), noise=noise, seed=2000 + t)
            sf = refo.Subframe(); sf.clear()
            L["phy"].get_dl_subframe_and_ce(L["i"], L["q"], 0, subfr, cell, 1, sf)
            err, rcfi, nsym, ral = L["phy"].pdcch_decode2(sf, cell, 1)
            r = port.pdcch_channel_decode(sf.rx_symb_re.copy(), sf.rx_symb_im.copy(), sf.rx_ce_re[0].copy(), sf.rx_ce_im[0].copy(), nrb, cell, 1, subfr)
            tag = (t, nrb, cell, cfi, subfr, n_msg, noise)
            assert (r["err"], r["cfi"], r["N_symbs"], len(r["allocs"])) == (err, rcfi, nsym, len(ral)), tag
            for a, want in zip(r["allocs"], ral):
                assert np.array_equal(alloc_row(a), alloc_row(want)), tag
            n_dci += len(ral)
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)
    assert n_dci >= 24


def test_product_dci_unpack_against_live_reference_random_payloads(refo):
    """lteb200_pdcch_unpack (host code) against the reference's dci_1a_unpack / dci_1c_unpack (liblte_phy.cc:13273-13605) on fresh random payloads:
    every bandwidth, SI / P / RA RNTIs (the only ones the blind decode hands over), 1 / 2 / 4 antenna ports, both formats -- error verdict and every allocation field (RIV decoding,
    distributed-VRB interleaver of format 1C, TBS look-ups, modulation, redundancy version, transmission mode)."""
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    import ctypes as C
    import sys
    import openlte_b200
    from openlte_b200.lib import PdcchRaw, Dci
    lib = openlte_b200.load()
    rng = np.random.default_rng(1327)
    n_ok = n_err = 0
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)     # the reference prints a line per "format 0" DCI
    try:
        for N in (6, 15, 25, 50, 75, 100):
            for t in range(150):
                bits = rng.integers(0, 2, 32).astype(np.uint8)
                if t % 2:
                    bits[0] = 1                                    # format 1A flag set: otherwise the reference rejects half of the draws at once
                rnti = int(rng.choice([0xFFFF, 0xFFFE, int(rng.integers(1, 0x3D))]))      # what the blind decode hands over (liblte_phy.cc:4919-4945): SI, P, RA
                n_ant = int(rng.choice([1, 2, 4]))
                for fmt, fn in ((0, refo.dci_1a_unpack), (1, refo.dci_1c_unpack)):
                    err, want = fn(bits, rnti, N, n_ant)
                    raw = PdcchRaw(); raw.cfi = 2; raw.N_symbs = 2
                    raw.found[fmt] = 1; raw.rnti[fmt] = rnti
                    for i in range(32):
                        raw.bits[fmt][i] = int(bits[i])
                    allocs = (Dci * 6)(); n = C.c_int(0)
                    rc = lib.pdcch_unpack(C.byref(raw), N, n_ant, allocs, C.byref(n))
                    assert (rc != 0) == (err != 0), (N, fmt, hex(rnti), n_ant, rc, err)
                    if err == 0:
                        assert n.value == 1 and np.array_equal(alloc_row(allocs[0].as_dict()), alloc_row(want)), (N, fmt, hex(rnti), n_ant)
                        n_ok += 1
                    else:
                        n_err += 1
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)
    assert n_ok > 300 and n_err > 100, (n_ok, n_err)

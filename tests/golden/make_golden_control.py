"""Control-channel fixtures (PBCH, PCFICH/PDCCH) from the REFERENCE ITSELF: its transmit functions build the subframe, its
receive front end produces grid + channel estimate, and its decoders give the expected results.

    python tests/golden/make_golden_control.py        (build container only: needs oracle/_ref and /root/reference)

Stored per case: the reference's rx grid [16][nsc] and channel estimates [ports][16][nsc] (float32), the configuration, and the
reference's decode result.  The -m gpu tests feed the grid to the CUDA path and must reproduce the result exactly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
from util import ref_loopback  # noqa: E402

FS = {100: 30.72, 75: 30.72, 50: 15.36, 25: 7.68, 15: 3.84, 6: 1.92}


def alloc_rows(allocs):
    out = np.zeros((len(allocs), 228), np.int32)
    for r, a in zip(out, allocs):
        r[:8] = [a["rnti"], a["tbs"], a["mcs"], a["N_prb"], a["rv_idx"], a["mod_type"], a["tx_mode"], a["chan_type"]]
        r[8:8 + len(a["prb"])] = a["prb"]; r[118:118 + len(a["prb1"])] = a["prb1"]
    return out


def bch_cases():
    rng = np.random.default_rng(11)
    out = {}
    cases = [(100, 0, 1, 0, 0.0), (25, 17, 2, 1, 0.02), (50, 301, 4, 2, 0.02), (6, 77, 1, 3, 0.05), (6, 150, 2, 7, 0.02), (15, 503, 4, 5, 0.01),
             (15, 9, 1, 2, 400.0)]          # last one: too noisy, must fail in the same way
    for n, (nrb, cell, nant, sfn, noise) in enumerate(cases):
        bits = rng.integers(0, 2, 24).astype(np.uint8)
        L = ref_loopback(ref, 0, [], N_id_cell=cell, N_rb_dl=nrb, fs=FS[nrb], N_ant=nant, bch=dict(bits=bits, sfn=sfn), noise=noise, seed=cell)
        phy = L["phy"]; sf = ref.Subframe(); sf.clear()
        phy.get_dl_subframe_and_ce(L["i"], L["q"], 0, 0, cell, 4, sf)
        err, na, ob, off = phy.bch_decode(sf, cell)
        nsc = 12 * nrb
        out[f"cfg_{n}"] = np.array([nrb, cell, nant, sfn, err, na, off], np.int32)
        out[f"tx_bits_{n}"] = bits
        out[f"rx_bits_{n}"] = np.asarray(ob[:24], np.uint8) if err == 0 else np.zeros(24, np.uint8)
        out[f"rx_re_{n}"] = sf.rx_symb_re[:, :nsc].copy(); out[f"rx_im_{n}"] = sf.rx_symb_im[:, :nsc].copy()
        out[f"ce_re_{n}"] = sf.rx_ce_re[:, :, :nsc].copy(); out[f"ce_im_{n}"] = sf.rx_ce_im[:, :, :nsc].copy()
        print("bch", nrb, cell, nant, sfn, "->", err, na, off)
    out["n_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(HERE, "control_bch_ref.npz"), **out)


def pdcch_cases():
    out = {}
    SI, P = 0xFFFF, 0xFFFE
    cases = [(100, 0, 2, 1, [dict(rnti=SI, mcs=5, N_prb=4, prb0=0)], 0.0),
             (100, 17, 3, 4, [dict(rnti=SI, mcs=9, N_prb=8, prb0=10), dict(rnti=P, mcs=2, N_prb=3, prb0=40), dict(rnti=7, mcs=1, N_prb=2, prb0=60)], 0.05),
             (50, 301, 1, 7, [dict(rnti=P, mcs=3, N_prb=5, prb0=3)], 0.05),
             (25, 77, 3, 5, [dict(rnti=SI, mcs=4, N_prb=4, prb0=2)], 0.02),
             (25, 0, 1, 1, [dict(rnti=SI, mcs=2, N_prb=2, prb0=1)], 0.0),       # 3 CCEs only: level-4 candidate runs past the control region
             (15, 1, 1, 1, [dict(rnti=SI, mcs=2, N_prb=2, prb0=1)], 0.0),
             (6, 0, 2, 1, [dict(rnti=60, mcs=2, N_prb=2, prb0=1)], 0.0),
             (100, 5, 2, 6, [], 0.05),                                          # PCFICH only
             (100, 300, 2, 8, [dict(rnti=SI, mcs=5, N_prb=4, prb0=0)], 0.6),    # noisy
             (15, 500, 3, 2, [dict(rnti=SI, mcs=2, N_prb=3, prb0=1)], 0.02)]    # PCFICH fails in the reference
    for n, (nrb, cell, cfi, subfr, allocs, noise) in enumerate(cases):
        L = ref_loopback(ref, subfr, [], N_id_cell=cell, N_rb_dl=nrb, fs=FS[nrb], N_ant=1, pdcch=dict(cfi=cfi, allocs=allocs), noise=noise, seed=cell)
        phy = L["phy"]; sf = ref.Subframe(); sf.clear()
        phy.get_dl_subframe_and_ce(L["i"], L["q"], 0, subfr, cell, 1, sf)
        err, rcfi, nsym, ral = phy.pdcch_decode2(sf, cell, 1)
        nsc = 12 * nrb
        out[f"cfg_{n}"] = np.array([nrb, cell, cfi, subfr, err, rcfi, nsym, len(ral)], np.int32)
        out[f"allocs_{n}"] = alloc_rows(ral)
        out[f"rx_re_{n}"] = sf.rx_symb_re[:, :nsc].copy(); out[f"rx_im_{n}"] = sf.rx_symb_im[:, :nsc].copy()
        out[f"ce_re_{n}"] = sf.rx_ce_re[0, :, :nsc].copy(); out[f"ce_im_{n}"] = sf.rx_ce_im[0, :, :nsc].copy()
        print("pdcch", nrb, cell, cfi, subfr, "->", err, rcfi, nsym, [(a["rnti"], a["tbs"], a["N_prb"]) for a in ral])
    out["n_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(HERE, "control_pdcch_ref.npz"), **out)


def pieces():
    """rate_unmatch_conv + viterbi_decode on random integer-valued soft bits, DCI unpacking on random payloads, TBS columns."""
    rng = np.random.default_rng(3)
    phy = ref.RefPhy()
    out = {}
    n = 0
    for t in range(60):
        N_c = int(rng.integers(20, 60)); E = int(rng.choice([72, 144, 288, 576, 1920, 100]))
        e = np.round(rng.standard_normal(E) * 40).astype(np.float32)
        if t % 3 == 0:
            e[rng.integers(0, E, E // 3)] = 10000.0
        d = phy.rate_unmatch_conv(e, N_c)
        out[f"cc_e_{n}"] = e; out[f"cc_nc_{n}"] = np.int32(N_c); out[f"cc_d_{n}"] = d; out[f"cc_c_{n}"] = phy.viterbi_decode(d)
        n += 1
    out["n_cc"] = np.int32(n)
    rows = []
    for N in (6, 15, 25, 50, 75, 100):
        for t in range(40):
            bits = rng.integers(0, 2, 32).astype(np.uint8); rnti = int(rng.choice([0xFFFF, 0xFFFE, 5]))
            for fmt, fn in ((0, ref.dci_1a_unpack), (1, ref.dci_1c_unpack)):
                err, a = fn(bits, rnti, N, 1)
                rows.append(np.concatenate([[N, fmt, rnti, err], bits, alloc_rows([a])[0]]).astype(np.int32))
    out["dci_rows"] = np.stack(rows)
    out["tbs_1a"] = np.array([[ref.tbs_71721(i, 2), ref.tbs_71721(i, 3)] for i in range(27)], np.int32)
    out["tbs_1c"] = np.array([ref.tbs_71723(i) for i in range(32)], np.int32)
    np.savez_compressed(os.path.join(HERE, "control_pieces_ref.npz"), **out)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "scan"):
    # the reference prints a line for every DCI 1A whose first bit says "format 0": keep the terminal readable
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        pieces()
    finally:
        os.dup2(saved, 1)
    bch_cases()
    pdcch_cases()


# ------------------------------------------------------------------------------------------------------------ scan loop
def scan_subframe(phy, i, q, frame_start, sn, cell):
    """One iteration of the reference's scan loop (LTE_fdd_dl_fs_samp_buf.cc:491-515): front end, PDCCH decode, PDSCH decode of
    every allocation found.  Returns the row the fixtures store."""
    sf = ref.Subframe(); sf.clear()
    phy.get_dl_subframe_and_ce(i, q, frame_start & 0xFFFFFFFF, sn, cell, 1, sf)
    err, cfi, nsym, allocs = phy.pdcch_decode2(sf, cell, 1)
    tbs_bits = []
    oks = []
    if err == 0:
        for a in allocs:
            if a["N_prb"] < 1 or a["tbs"] < 1 or a["prb"] != a["prb1"]:
                oks.append(0); tbs_bits.append(np.zeros(0, np.uint8)); continue
            e2, bits = phy.pdsch_decode(sf, nsym, cell, 1, a["mod_type"], a["prb"], a["tbs"], a["rnti"], rv=a["rv_idx"], tx_mode=a["tx_mode"])
            oks.append(1 if e2 == 0 else 0); tbs_bits.append(bits if e2 == 0 else np.zeros(0, np.uint8))
    return err, cfi, nsym, allocs, oks, tbs_bits


def store_scan(out, tag, rows):
    out[f"{tag}_n"] = np.int32(len(rows))
    for k, (err, cfi, nsym, allocs, oks, bits) in enumerate(rows):
        out[f"{tag}_hdr_{k}"] = np.array([err, cfi, nsym, len(allocs)], np.int32)
        out[f"{tag}_allocs_{k}"] = alloc_rows(allocs)
        out[f"{tag}_ok_{k}"] = np.array(oks, np.int32)
        for a, b in enumerate(bits):
            out[f"{tag}_bits_{k}_{a}"] = np.asarray(b, np.uint8)


def scan_cases():
    out = {}
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
    # A: the shipped capture, subframes 1..9 of the frame that starts 15 samples before the buffer
    z = np.load(os.path.join(HERE, "capture_cfg1.npz"))
    iq = z["iq"]; i = np.ascontiguousarray(iq[:, 0].astype(np.float32)); q = np.ascontiguousarray(iq[:, 1].astype(np.float32))
    fs0 = int(z["frame_start_idx"]) - 307200
    phy = ref.RefPhy(fs=30.72, N_id_cell=0, N_ant=1, N_rb_dl=100)
    os.dup2(devnull, 1)
    try:
        rows = [scan_subframe(phy, i, q, fs0, sn, 0) for sn in range(1, 10)]
    finally:
        os.dup2(saved, 1)
    store_scan(out, "cap", rows)
    out["cap_frame_start"] = np.int64(fs0)
    print("capture scan:", [(r[0], r[1], len(r[3]), r[4]) for r in rows])
    # B: synthetic 10 MHz stream from the reference's transmitter: DCI 1A + the PDSCH it announces, four subframes + CRS-only tail
    rng = np.random.default_rng(21)
    nrb, cell, fs, n_s = 50, 23, 15.36, 15360
    phy = ref.RefPhy(fs=fs, N_id_cell=cell, N_ant=1, N_rb_dl=nrb)
    plan = {1: (2, [dict(rnti=0xFFFF, mcs=5, N_prb=6, prb0=3)]),
            2: (3, [dict(rnti=0xFFFE, mcs=9, N_prb=8, prb0=10), dict(rnti=7, mcs=1, N_prb=2, prb0=30)]),
            3: (1, []),
            4: (2, [dict(rnti=0xFFFF, mcs=2, N_prb=3, prb0=40)])}
    total = 6 * n_s + 4096
    si = np.zeros(total, np.float32); sq = np.zeros(total, np.float32)
    tx_bits = {}
    for sn in range(1, 6):
        sf = ref.Subframe(); sf.clear(); sf.set_num(sn)
        phy.map_crs(sf, cell, 1)
        if sn in plan:
            cfi, allocs = plan[sn]
            for a in allocs:
                tbs = ref.tbs_71721(a["mcs"], 3)
                tb = rng.integers(0, 2, tbs).astype(np.uint8); tx_bits[(sn, a["rnti"])] = tb
                nsym = cfi
                assert phy.pdsch_encode(sf, nsym, cell, 1, 1, list(range(a["prb0"], a["prb0"] + a["N_prb"])), tbs, a["rnti"], tb) == 0
            assert phy.pdcch_encode(sf, cell, 1, cfi, allocs) == 0
        a_i, a_q = phy.create_dl_subframe(sf, 0, n_s)
        si[sn * n_s:(sn + 1) * n_s] = a_i; sq[sn * n_s:(sn + 1) * n_s] = a_q
    z_ = (si + 1j * sq) * (0.8 * np.exp(0.4j)) + 0.01 * (rng.standard_normal(total) + 1j * rng.standard_normal(total))
    sc = 100.0 / max(np.abs(z_.real).max(), np.abs(z_.imag).max())
    i8 = np.stack([np.rint(z_.real * sc), np.rint(z_.imag * sc)], -1).astype(np.int8)
    i = np.ascontiguousarray(i8[:, 0].astype(np.float32)); q = np.ascontiguousarray(i8[:, 1].astype(np.float32))
    os.dup2(devnull, 1)
    try:
        rows = [scan_subframe(phy, i, q, 0, sn, cell) for sn in range(1, 5)]
    finally:
        os.dup2(saved, 1)
    store_scan(out, "syn", rows)
    out["syn_iq"] = i8; out["syn_cfg"] = np.array([nrb, cell, 3], np.int32)      # N_rb_dl, cell, fs_idx
    print("synthetic scan:", [(r[0], r[1], len(r[3]), r[4]) for r in rows])
    for k, sn in enumerate(range(1, 5)):
        for a, al in enumerate(rows[k][3]):
            if rows[k][4][a] and (sn, al["rnti"]) in tx_bits:
                assert np.array_equal(rows[k][5][a], tx_bits[(sn, al["rnti"])]), (sn, al["rnti"])
    np.savez_compressed(os.path.join(HERE, "control_scan_ref.npz"), **out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "scan":
    scan_cases()

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


if __name__ == "__main__":
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

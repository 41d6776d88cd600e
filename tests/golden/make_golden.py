"""Generates the fixtures under tests/golden/ from the REFERENCE ITSELF (oracle/_ref = /root/reference's
liblte_phy.cc compiled unmodified).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; the -m gpu tests read only these committed fixtures (and oracle/_ref/*.so,
which travels with the snapshot).  Every array is float32/int8/uint8 to keep the files small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
from util import ref_loopback  # noqa: E402

FS_IDX = {1.92: 0, 3.84: 1, 7.68: 2, 15.36: 3, 30.72: 4}
QM = {0: 1, 1: 2, 2: 4, 3: 6}


def loopback_case(name, fs, N_rb_dl, cell, subfr, mod, prb, tbs, rnti, N_ant_ce=1, gain=0.7 * np.exp(0.9j), noise=0.002, cfi=2, seed=0,
                  frame_start=0, N_ant_tx=1):
    lb = ref_loopback(ref, subfr, [dict(mod=mod, prb=prb, tbs=tbs, rnti=rnti)], N_id_cell=cell, cfi=cfi, N_rb_dl=N_rb_dl, fs=fs, seed=seed,
                      gain=gain, noise=noise, frame_start=frame_start, N_ant=N_ant_tx)
    phy = lb["phy"]
    sf = ref.Subframe()
    phy.get_dl_subframe_and_ce(lb["i"], lb["q"], frame_start, subfr, cell, N_ant_ce, sf)
    rx_re, rx_im = sf.rx_symb_re.copy(), sf.rx_symb_im.copy()
    ce_re, ce_im = sf.rx_ce_re.copy(), sf.rx_ce_im.copy()
    # PDSCH decode uses as many ports as the transmitter has
    sf1 = ref.Subframe()
    phy.get_dl_subframe_and_ce(lb["i"], lb["q"], frame_start, subfr, cell, N_ant_tx, sf1)
    err, bits = phy.pdsch_decode(sf1, cfi, cell, N_ant_tx, mod, prb, tbs, rnti)
    from oracle import port
    n_re = port.pdsch_re_indices(N_rb_dl, cell, N_ant_tx, subfr, cfi, prb).size
    G = n_re * QM[mod]
    descr = phy.pdsch_scratch(1, G, np.float32)
    # trim the sample arrays to what the receiver touches
    n_keep = frame_start + (subfr + 1) * lb["n_sf"] + 3 * (128 << FS_IDX[fs]) * 2
    np.savez_compressed(os.path.join(HERE, name), i=lb["i"][:n_keep], q=lb["q"][:n_keep], fs_idx=FS_IDX[fs], N_rb_dl=N_rb_dl, N_id_cell=cell,
                        subfr=subfr, n_sf=lb["n_sf"], frame_start=frame_start, mod=mod, prb=np.array(prb, np.int32), tbs=tbs, rnti=rnti, cfi=cfi,
                        N_ant_ce=N_ant_ce, N_ant_tx=N_ant_tx, tb=lb["tb"][0], ref_rx_re=rx_re, ref_rx_im=rx_im, ref_ce_re=ce_re[:, :14], ref_ce_im=ce_im[:, :14],
                        ref_descramb=descr, ref_crc_ok=int(err == 0), ref_tb=bits)
    print(name, "G", G, "ref err", err, "tb ok", np.array_equal(bits, lb["tb"][0]))


def turbo_cases():
    """A few reference-decoded code words per K (inputs + the reference's decoded bits and intermediates)."""
    phy = ref.RefPhy()
    rng = np.random.default_rng(42)
    out = {}
    for K in (40, 1024, 3584, 6144):
        for sigma in (0.0, 0.5, 0.8):
            c = rng.integers(0, 2, K).astype(np.uint8)
            d = phy.turbo_encode(c)
            x = (1.0 - 2.0 * d.astype(np.float32)).reshape(3, K + 4).T.reshape(-1)
            x = (x + sigma * rng.standard_normal(x.size)).astype(np.float32)
            if sigma == 0.5:
                x[1::21] = 10000.0
            key = f"K{K}_s{int(sigma * 10)}"
            out[key + "_c"] = c
            out[key + "_enc"] = d
            out[key + "_x"] = x.astype(np.float16).astype(np.float32)  # store exactly representable values, half the bytes
            out[key + "_dec"] = phy.turbo_decode(out[key + "_x"])
            out[key + "_act1"] = phy.turbo_intermediate(0, K)
    np.savez_compressed(os.path.join(HERE, "turbo_ref.npz"), **out)
    print("turbo_ref.npz", len(out))


def ratematch_cases():
    phy = ref.RefPhy()
    rng = np.random.default_rng(43)
    out = {}
    for K in (40, 1024, 6144):
        D = K + 4
        for rv in range(4):
            for E in (3 * D, 6 * D + 8, 2 * D + 3):
                d = rng.integers(0, 2, 3 * D).astype(np.uint8)
                e = phy.rate_match_turbo(d, E, rv=rv)
                llr = (rng.standard_normal(E) * 4).astype(np.float16).astype(np.float32)
                llr[::19] = 10000.0
                key = f"K{K}_rv{rv}_E{E}"
                out[key + "_d"] = d; out[key + "_e"] = e; out[key + "_llr"] = llr.astype(np.float16)
                out[key + "_un"] = phy.rate_unmatch_turbo(llr, D, rv=rv)
    np.savez_compressed(os.path.join(HERE, "ratematch_ref.npz"), **out)
    print("ratematch_ref.npz", len(out))


def sequence_cases():
    out = {}
    for ci in (0, 1, 0x3FFF0123, (0xFFFF << 14) | (5 << 9), 12345678):
        out[f"prs_{ci}"] = ref.prs_c(ci, 2000).astype(np.uint8)
    rng = np.random.default_rng(44)
    for n in (24, 144, 6120):
        b = rng.integers(0, 2, n).astype(np.uint8)
        out[f"crc_in_{n}"] = b
        out[f"crc24a_{n}"] = ref.crc(b, 0); out[f"crc24b_{n}"] = ref.crc(b, 1); out[f"crc16_{n}"] = ref.crc(b, 2)
    for (ns, l, cell) in ((0, 0, 0), (19, 4, 503), (7, 1, 77)):
        re, im = ref.crs(ns, l, cell)
        out[f"crs_{ns}_{l}_{cell}"] = np.stack([re, im])
    for mod in (0, 1, 2, 3):
        re = (rng.standard_normal(600) * 0.8).astype(np.float32); im = (rng.standard_normal(600) * 0.8).astype(np.float32)
        re[:5] = 0; im[3:8] = 0
        out[f"demap_in_{mod}"] = np.stack([re, im]); out[f"demap_out_{mod}"] = ref.mod_demap(re, im, mod)
    np.savez_compressed(os.path.join(HERE, "sequences_ref.npz"), **out)
    print("sequences_ref.npz", len(out))


def capture_case():
    """BASELINE config 1: the first 340,000 samples of the reference's shipped capture (20 MHz, cell 0) and what the
    reference finds in them: coarse timing (20 slots), PSS, SSS, and SIB1 of the frame that starts 15 samples before the
    buffer (sfn even) -- PDCCH decode supplies the allocation, PDSCH decode the 144 transport-block bits."""
    raw = np.fromfile("/root/reference/Test captures/LTE_test_file_int8.bin", dtype=np.int8)
    N = 345000          # through subframe 0 of the frame at 307185 plus the estimator's two look-ahead symbols
    iq = raw[:2 * N].reshape(N, 2).copy()
    i = np.ascontiguousarray(iq[:, 0].astype(np.float32)); q = np.ascontiguousarray(iq[:, 1].astype(np.float32))
    phy = ref.RefPhy(fs=30.72, N_id_cell=0xFFFF, N_ant=4, N_rb_dl=6)
    err, fo, ss, n = phy.find_coarse_timing(i, q, 20)
    abs_corr = phy.timing_abs_corr(15360)
    err, ss2, nid2, ps, th, f = phy.find_pss(i, q, ss[0])
    err, ss3, nid1, fsi = phy.find_sss(i, q, nid2, ss2, th)
    # MIB from subframe 0 of the detected frame (LTE_fdd_dl_fs_samp_buf.cc:378-386: four-port estimate, N_rb_dl still 6)
    sf0 = ref.Subframe()
    phy.get_dl_subframe_and_ce(i, q, fsi, 0, 3 * nid1 + nid2, 4, sf0)
    berr, mib_ant, mib_bits, mib_off = phy.bch_decode(sf0, 3 * nid1 + nid2)
    assert berr == 0
    phy.update_n_rb_dl(100)
    sf = ref.Subframe()
    fs0 = (fsi - 307200) & 0xFFFFFFFF
    phy.get_dl_subframe_and_ce(i, q, fs0, 5, 3 * nid1 + nid2, 1, sf)
    err, cfi, allocs = phy.pdcch_decode(sf, 0, 1)
    a = allocs[0]
    err, bits = phy.pdsch_decode(sf, cfi, 0, 1, a["mod_type"], a["prb"], a["tbs"], a["rnti"], rv=a["rv_idx"], tx_mode=a["tx_mode"])
    assert err == 0
    np.savez_compressed(os.path.join(HERE, "capture_cfg1.npz"), iq=iq, coarse_n=n, coarse_fo=fo, coarse_ss=ss, pss_ss=ss2, N_id_2=nid2, pss_symb=ps,
                        pss_thresh=np.float32(th), pss_fo=np.float32(f), sss_ss=ss3, N_id_1=nid1, frame_start_idx=fsi, sib1_frame_start=np.uint32(fs0),
                        cfi=cfi, rnti=a["rnti"], tbs=a["tbs"], mod=a["mod_type"], rv=a["rv_idx"], prb=np.array(a["prb"], np.int32), sib1_bits=bits,
                        sib1_rx_re=sf.rx_symb_re.copy(), sib1_rx_im=sf.rx_symb_im.copy(), sib1_ce_re=sf.rx_ce_re[0, :14].copy(),
                        sib1_ce_im=sf.rx_ce_im[0, :14].copy(), mib_bits=np.asarray(mib_bits[:24], np.uint8), mib_N_ant=mib_ant, mib_offset=mib_off)
    print("capture_cfg1.npz", n, ss[0][0], nid2, ps, th, nid1, fsi, "".join(map(str, bits)), "MIB", "".join(map(str, mib_bits[:24])), mib_ant, mib_off)


if __name__ == "__main__":
    capture_case()
    if len(sys.argv) > 1 and sys.argv[1] == "capture":
        sys.exit(0)
    loopback_case("loop_64qam_c0_sf1.npz", 30.72, 100, 0, 1, 3, list(range(12)), 2216, 0x1234)
    loopback_case("loop_qpsk_c77_sf5.npz", 30.72, 100, 77, 5, 1, list(range(40, 76)), 3240, 0x0A0B, N_ant_ce=2, seed=1)
    loopback_case("loop_16qam_c150_sf0.npz", 30.72, 100, 150, 0, 2, list(range(41, 59)), 3112, 0x7FFF, seed=2)
    loopback_case("loop_qpsk_5mhz_c301_sf3.npz", 7.68, 25, 301, 3, 1, list(range(2, 20)), 1544, 0x0042, N_ant_ce=4, seed=3, frame_start=137)
    loopback_case("loop_2ant_16qam_c21_sf2.npz", 30.72, 100, 21, 2, 2, list(range(10, 28)), 2216, 0x0F0F, N_ant_ce=2, N_ant_tx=2, seed=4)
    loopback_case("loop_4ant_qpsk_10mhz_c400_sf5.npz", 15.36, 50, 400, 5, 1, list(range(20, 50)), 2216, 0x1111, N_ant_ce=4, N_ant_tx=4, seed=5)
    loopback_case("loop_2ant_64qam_c5_sf7.npz", 30.72, 100, 5, 7, 3, list(range(60, 72)), 2216, 0x2222, N_ant_ce=2, N_ant_tx=2, seed=6)
    turbo_cases()
    ratematch_cases()
    sequence_cases()

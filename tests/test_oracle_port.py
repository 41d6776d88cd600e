"""The oracle is pinned here (CPU only): every function of the restatement in oracle/port/ is compared with the
fixtures generated from the reference itself (tests/golden/make_golden.py) and, where oracle/_ref is present, with
the reference directly on fresh random inputs."""
import os

import numpy as np
import pytest

from util import noisy_codeword

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def test_qpp_table_is_the_spec_and_pi_overflows_like_the_reference(port):
    # every row of the table is a permutation in exact arithmetic ...
    from oracle.port import qpp_pi
    bad = []
    import re
    rows = re.findall(r"\{\s*(\d+),\s*(\d+),\s*(\d+)\}", open(os.path.join(os.path.dirname(GOLD), "..", "openlte_b200", "csrc", "qpp_table.h")).read())
    assert len(rows) == 188
    for K, f1, f2 in [(int(a), int(b), int(c)) for a, b, c in rows]:
        exact = [(f1 * i + f2 * i * i) % K for i in range(K)]
        assert len(set(exact)) == K
        got = qpp_pi(K)
        if not np.array_equal(got, np.array(exact, np.uint32)):
            bad.append((K, len(set(got.tolist()))))
    # ... and exactly 20 sizes differ from it under uint32 wrap-around, K = 6144 keeping 3955 distinct indices
    assert len(bad) == 20 and (6144, 3955) in bad and (3584, 3576) in bad and (5248, 2500) in bad


def test_turbo_against_reference_fixtures(port):
    g = gold("turbo_ref.npz")
    for K in (40, 1024, 3584, 6144):
        for s in (0, 5, 8):
            key = f"K{K}_s{s}"
            assert np.array_equal(port.turbo_encode(g[key + "_c"]), g[key + "_enc"])
            dec, dbg = port.turbo_decode_ref(g[key + "_x"], True)
            assert np.array_equal(dec, g[key + "_dec"]), key
            assert np.array_equal(dbg[0], g[key + "_act1"]), key
            if s == 0:
                assert np.array_equal(dec, g[key + "_c"])
                assert np.array_equal(port.turbo_decode_mlm(g[key + "_x"], 8), g[key + "_c"])


def test_ratematch_against_reference_fixtures(port):
    g = gold("ratematch_ref.npz")
    keys = sorted({k.rsplit("_", 1)[0] for k in g})
    assert len(keys) == 36
    for key in keys:
        K = int(key.split("_")[0][1:]); rv = int(key.split("_")[1][2:]); E = int(key.split("_")[2][1:])
        assert np.array_equal(port.rate_match_turbo(g[key + "_d"], E, rv=rv), g[key + "_e"]), key
        assert np.array_equal(port.rate_unmatch_turbo(g[key + "_llr"].astype(np.float32), K + 4, rv=rv), g[key + "_un"]), key


def test_sequences_crc_demap_against_reference_fixtures(port):
    g = gold("sequences_ref.npz")
    for ci in (0, 1, 0x3FFF0123, (0xFFFF << 14) | (5 << 9), 12345678):
        assert np.array_equal(port.prs_c(ci, 2000), g[f"prs_{ci}"])
    for n in (24, 144, 6120):
        b = g[f"crc_in_{n}"]
        as_int = lambda bits: int("".join(map(str, bits)), 2)
        assert port.crc(b, 0x1864CFB, 24) == as_int(g[f"crc24a_{n}"])
        assert port.crc(b, 0x1800063, 24) == as_int(g[f"crc24b_{n}"])
        assert port.crc(b, 0x11021, 16) == as_int(g[f"crc16_{n}"])
    for (ns, l, cell) in ((0, 0, 0), (19, 4, 503), (7, 1, 77)):
        re, im = port.crs(ns, l, cell)
        assert np.array_equal(np.stack([re, im]), g[f"crs_{ns}_{l}_{cell}"])
    for mod in (0, 1, 2, 3):
        re, im = g[f"demap_in_{mod}"]
        assert np.array_equal(port.demap(re, im, mod), g[f"demap_out_{mod}"])
    assert port.segmentation_params(144) == (168, 0, 1)
    assert port.segmentation_params(6120) == (6144, 0, 1)
    assert port.segmentation_params(100)[:2] == (128, 4)


@pytest.mark.parametrize("name", ["loop_64qam_c0_sf1.npz", "loop_qpsk_c77_sf5.npz", "loop_16qam_c150_sf0.npz", "loop_qpsk_5mhz_c301_sf3.npz"])
def test_frontend_against_reference_fixtures(port, name):
    c = gold(name)
    nsc = 12 * int(c["N_rb_dl"]); n_ant = int(c["N_ant_ce"])
    # channel estimate from the reference's own grid: bit-identical (same libm, same operation order)
    ce_re, ce_im = port.chan_est(c["ref_rx_re"], c["ref_rx_im"], int(c["N_rb_dl"]), int(c["subfr"]), int(c["N_id_cell"]), n_ant)
    assert np.array_equal(ce_re[:n_ant, :14, :nsc], c["ref_ce_re"][:n_ant, :, :nsc])
    assert np.array_equal(ce_im[:n_ant, :14, :nsc], c["ref_ce_im"][:n_ant, :, :nsc])
    # OFDM demod: the port's DFT is double precision, the reference's is single: tolerance 2e-6 of the grid maximum
    rx_re, rx_im, _, _ = port.get_dl_subframe_and_ce(c["i"], c["q"], int(c["fs_idx"]), int(c["N_rb_dl"]), int(c["frame_start"]), int(c["subfr"]),
                                                     int(c["N_id_cell"]), 1)
    sc = np.abs(c["ref_rx_re"]).max()
    assert np.abs(rx_re - c["ref_rx_re"]).max() <= 2e-6 * sc and np.abs(rx_im - c["ref_rx_im"]).max() <= 2e-6 * sc
    # RE extraction + ZF + demap + descramble: exactly the reference's pdsch_descramb_bits
    ce1_re, ce1_im = port.chan_est(c["ref_rx_re"], c["ref_rx_im"], int(c["N_rb_dl"]), int(c["subfr"]), int(c["N_id_cell"]), 1)
    idx = port.pdsch_re_indices(int(c["N_rb_dl"]), int(c["N_id_cell"]), 1, int(c["subfr"]), int(c["cfi"]), c["prb"])
    # grid rows are 1200 wide in the reference struct; indices are L*1200+sc
    xr, xi = port.zf_1ant(c["ref_rx_re"].reshape(-1)[idx], c["ref_rx_im"].reshape(-1)[idx], ce1_re[0].reshape(-1)[idx], ce1_im[0].reshape(-1)[idx])
    soft = port.demap(xr, xi, int(c["mod"])).astype(np.float32)
    cseq = port.prs_c((int(c["rnti"]) << 14) | (int(c["subfr"]) << 9) | int(c["N_id_cell"]), soft.size)
    assert np.array_equal(soft * (1 - 2 * cseq.astype(np.float32)), c["ref_descramb"])
    rc, bits = port.pdsch_decode(c["ref_rx_re"], c["ref_rx_im"], ce1_re[0], ce1_im[0], int(c["N_rb_dl"]), int(c["N_id_cell"]), int(c["subfr"]),
                                 int(c["cfi"]), int(c["mod"]), c["prb"], int(c["tbs"]), int(c["rnti"]))
    assert (rc == 0) == bool(c["ref_crc_ok"])
    if rc == 0:
        assert np.array_equal(bits, c["ref_tb"]) and np.array_equal(bits, c["tb"])
    rc, bits = port.pdsch_decode(c["ref_rx_re"], c["ref_rx_im"], ce1_re[0], ce1_im[0], int(c["N_rb_dl"]), int(c["N_id_cell"]), int(c["subfr"]),
                                 int(c["cfi"]), int(c["mod"]), c["prb"], int(c["tbs"]), int(c["rnti"]), turbo_mode=1)
    assert rc == 0 and np.array_equal(bits, c["tb"])      # max-log-MAP also decodes the punctured case the reference fails (H11)


def test_port_against_live_reference(port, refo):
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    phy = refo.RefPhy()
    rng = np.random.default_rng(11)
    for K in (48, 512, 5248, 6144):
        for sigma in (0.0, 0.4, 0.9):
            c, x = noisy_codeword(port, K, sigma, rng, sentinel_every=13 if sigma == 0.4 else 0)
            assert np.array_equal(port.turbo_decode_ref(x), phy.turbo_decode(x))
        D = K + 4
        for rv in (0, 3):
            e = (rng.standard_normal(3 * D + 77) * 2).astype(np.float32)
            assert np.array_equal(port.rate_unmatch_turbo(e, D, rv=rv), phy.rate_unmatch_turbo(e, D, rv=rv))


def test_port_against_live_reference_every_block_size(port, refo):
    """All 188 block sizes of 36.212 table 5.1.3-3, one noisy code word each (noise level cycling through clean, waterfall and
    hopeless, sentinels on every third size): encoder, decoder (liblte_phy.cc:10541-10586, 10620-10845) and rate (un)matching
    (:11081-11490) of the port equal the live reference.  The GPU tests compare the kernels with the port for the same 188 sizes,
    so this closes the chain kernel == port == reference for every size, the 20 wrapped interleavers included."""
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    import re
    rows = re.findall(r"\{\s*(\d+),\s*(\d+),\s*(\d+)\}", open(os.path.join(os.path.dirname(GOLD), "..", "openlte_b200", "csrc", "qpp_table.h")).read())
    Ks = [int(a) for a, _, _ in rows]
    assert len(Ks) == 188
    phy = refo.RefPhy()
    rng = np.random.default_rng(188)
    for n, K in enumerate(Ks):
        sigma = (0.0, 0.45, 0.7, 1.2)[n % 4]
        c, x = noisy_codeword(port, K, sigma, rng, sentinel_every=11 if n % 3 == 0 else 0)
        assert np.array_equal(port.turbo_encode(c), phy.turbo_encode(c)), K
        assert np.array_equal(port.turbo_decode_ref(x), phy.turbo_decode(x)), K
        D = K + 4
        rv = n % 4
        e = (rng.standard_normal(2 * D + 3 * (n % 50) + 9) * 2).astype(np.float32)      # punctured (E < 3D) for every size
        assert np.array_equal(port.rate_unmatch_turbo(e, D, rv=rv), phy.rate_unmatch_turbo(e, D, rv=rv)), (K, rv)


@pytest.mark.parametrize("name", ["loop_2ant_16qam_c21_sf2.npz", "loop_4ant_qpsk_10mhz_c400_sf5.npz", "loop_2ant_64qam_c5_sf7.npz"])
def test_transmit_diversity_against_reference_fixtures(port, name):
    """2 / 4 antenna ports: SFBC combining + layer demapping + decode, from the reference's own grid."""
    c = gold(name)
    n_tx = int(c["N_ant_tx"])
    ce_re, ce_im = port.chan_est(c["ref_rx_re"], c["ref_rx_im"], int(c["N_rb_dl"]), int(c["subfr"]), int(c["N_id_cell"]), n_tx)
    assert np.array_equal(ce_re[:n_tx, :14], c["ref_ce_re"][:n_tx]) and np.array_equal(ce_im[:n_tx, :14], c["ref_ce_im"][:n_tx])
    rc, bits, soft = port.pdsch_decode_nant(c["ref_rx_re"], c["ref_rx_im"], ce_re, ce_im, int(c["N_rb_dl"]), int(c["N_id_cell"]), n_tx, int(c["subfr"]),
                                            int(c["cfi"]), int(c["mod"]), c["prb"], int(c["tbs"]), int(c["rnti"]))
    assert np.array_equal(soft, c["ref_descramb"])
    assert rc == 0 and np.array_equal(bits, c["ref_tb"]) and np.array_equal(bits, c["tb"])


def test_freq_shift_restatement(port):
    """fs.cc:696-713: rotation by exp(-j*2*pi*f*(i+1)/fs); the float product (i+1)*f limits the phase accuracy, so compare
    with a float64 rotation at 1e-3 and check the properties the scan relies on: power preserved, f then -f is the identity."""
    rng = np.random.default_rng(5)
    n, fs, f = 40000, 30720000, 1234.5
    i0 = rng.standard_normal(n).astype(np.float32); q0 = rng.standard_normal(n).astype(np.float32)
    i1, q1 = port.freq_shift(i0, q0, 100, n - 300, f, fs)
    assert np.array_equal(i1[:100], i0[:100]) and np.array_equal(q1[n - 200:], q0[n - 200:])
    k = np.arange(100, n - 200)
    want = (i0[k] + 1j * q0[k]) * np.exp(-2j * np.pi * f * (k + 1) / fs)
    got = i1[k] + 1j * q1[k]
    assert np.abs(got - want).max() < 1e-3
    assert np.allclose(np.abs(got), np.abs(i0[k] + 1j * q0[k]), rtol=1e-5, atol=1e-6)
    i2, q2 = port.freq_shift(i1, q1, 100, n - 300, -f, fs)
    assert np.abs(i2 - i0).max() < 1e-5 and np.abs(q2 - q0).max() < 1e-5


def test_front_end_and_pdsch_against_live_reference_random_configs(port, refo):
    """Randomised differential test against the reference itself (not fixtures): bandwidth, antenna ports, cell identity, subframe
    number, CFI, modulation, PRB set, transport-block size, RNTI, channel gain and noise are drawn afresh; the signal comes from the
    reference's own transmit functions.  Grid within the DFT tolerance (liblte_phy.cc:8593-8644); channel estimate (:5905-6200) and
    the PDSCH receive verdict + bits (:3690-3853) of the port, from the reference's grid, equal the reference's bit for bit --
    whether or not the block decodes."""
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    from util import ref_loopback
    rng = np.random.default_rng(2026)
    bands = [(30.72, 100, 4), (30.72, 75, 4), (15.36, 50, 3), (7.68, 25, 2), (3.84, 15, 1), (1.92, 6, 0)]
    seen_fail = seen_ok = 0
    for trial in range(42):
        fs, n_rb, fs_idx = bands[trial % len(bands)]
        n_ant = (1, 2, 4, 1, 1, 2, 1)[trial % 7]
        cell = int(rng.integers(0, 504)); subfr = int(rng.integers(0, 10)); cfi = int(rng.integers(1, 4)) if n_rb > 10 else int(rng.integers(2, 4))
        mod = int(rng.integers(1, 4))                                  # 1 QPSK, 2 16QAM, 3 64QAM
        n_prb = int(rng.integers(1, min(n_rb, 24) + 1))
        prb = np.sort(rng.choice(n_rb, n_prb, replace=False)).astype(np.int32)
        # code rate between 0.05 and 0.4 of the allocation's capacity (some trials will not decode: the verdicts must still agree)
        cap = n_prb * 120 * (2, 4, 6)[mod - 1]
        tbs = int(max(16, min(6120, int(cap * rng.uniform(0.05, 0.4)) // 8 * 8)))
        rnti = int(rng.integers(1, 65535))
        noise = float(rng.choice([0.0, 0.02, 0.08]))
        gain = complex(rng.uniform(0.5, 1.5) * np.exp(1j * rng.uniform(-np.pi, np.pi)))
        frame_start = int(rng.integers(0, 500))
        s = ref_loopback(refo, subfr, [dict(mod=mod, prb=prb, tbs=tbs, rnti=rnti)], N_id_cell=cell, cfi=cfi, N_rb_dl=n_rb, fs=fs, seed=trial,
                         gain=gain, noise=noise, frame_start=frame_start, N_ant=n_ant)
        phy = s["phy"]
        sf = refo.Subframe(); sf.clear()
        phy.get_dl_subframe_and_ce(s["i"], s["q"], frame_start, subfr, cell, n_ant, sf)
        rx_re, rx_im, ce_re, ce_im = port.get_dl_subframe_and_ce(s["i"], s["q"], fs_idx, n_rb, frame_start, subfr, cell, n_ant)
        tag = (trial, fs, n_rb, n_ant, cell, subfr, cfi, mod, n_prb, tbs)
        # the DFT is the one tolerance-held stage (port: double precision, reference: single): 2e-6 of the grid maximum ...
        g_re, g_im = sf.rx_symb_re.copy(), sf.rx_symb_im.copy()
        sc = max(np.abs(g_re).max(), np.abs(g_im).max())
        assert np.abs(rx_re - g_re).max() <= 2e-6 * sc and np.abs(rx_im - g_im).max() <= 2e-6 * sc, tag
        # ... everything after it is exact when it starts from the reference's own grid
        ce_re, ce_im = port.chan_est(g_re, g_im, n_rb, subfr, cell, n_ant)
        assert np.array_equal(ce_re[:n_ant, :14], sf.rx_ce_re[:n_ant, :14]) and np.array_equal(ce_im[:n_ant, :14], sf.rx_ce_im[:n_ant, :14]), tag
        err_ref, bits_ref = phy.pdsch_decode(sf, cfi, cell, n_ant, mod, prb, tbs, rnti)
        rc, bits, _ = port.pdsch_decode_nant(g_re, g_im, ce_re, ce_im, n_rb, cell, n_ant, subfr, cfi, mod, prb, tbs, rnti)
        assert (rc == 0) == (err_ref == 0), tag
        if rc == 0:
            assert np.array_equal(bits, bits_ref), tag
            seen_ok += 1
        else:
            seen_fail += 1
    assert seen_ok >= 15 and seen_fail >= 3, (seen_ok, seen_fail)          # the draw exercises both the decoding branch and DECODE_FAIL


def test_cell_search_against_live_reference(port, refo):
    """Cell search of the port against the reference itself on frames built by the reference's transmitter (CRS + PSS + SSS, random cell,
    timing offset, carrier offset, noise): coarse timing (liblte_phy.cc:5697-5852), PSS + fine timing (:5306-5510), SSS (:5578-5686).
    Every integer decision (peak count, symbol starts, N_id_2, PSS symbol, N_id_1, frame start) is equal; floats within 1e-4 relative
    (the port's DFT is double precision), the coarse frequency estimate within 1 Hz."""
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    rng = np.random.default_rng(77)
    cases = ((101, 2, 777, 30.72, 100, 4, 0.05, 0.0), (7, 1, 3000, 7.68, 25, 2, 0.05, 0.0), (55, 0, 123, 1.92, 6, 0, 0.1, 300.0), (167, 1, 2048, 15.36, 50, 3, 0.02, -900.0))
    for (nid1, nid2, off, fs, nrb, fsi_, noise, cfo) in cases:
        cell = 3 * nid1 + nid2
        phy = refo.RefPhy(fs=fs, N_id_cell=cell, N_ant=1, N_rb_dl=nrb)
        n_sf = 1920 << fsi_
        sig_i = np.zeros(off + 21 * n_sf + 8192, np.float32); sig_q = np.zeros_like(sig_i)
        for k in range(21):
            sf = refo.Subframe(); sf.clear(); sf.set_num(k % 10)
            phy.map_crs(sf, cell, 1)
            if k % 10 in (0, 5):
                phy.map_pss(sf, nid2, 1); phy.map_sss(sf, nid1, nid2, 1)
            a, b = phy.create_dl_subframe(sf, 0, n_sf)
            sig_i[off + k * n_sf: off + (k + 1) * n_sf] = a; sig_q[off + k * n_sf: off + (k + 1) * n_sf] = b
        z = sig_i + 1j * sig_q
        z = z * np.exp(2j * np.pi * cfo * np.arange(z.size) / (fs * 1e6))
        z = z + noise * np.abs(z).max() * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))
        i_s = np.ascontiguousarray(z.real.astype(np.float32)); q_s = np.ascontiguousarray(z.imag.astype(np.float32))
        tag = (nid1, nid2, off, fs)
        err, fo_r, ss_r, n_r = phy.find_coarse_timing(i_s, q_s, 20)
        fo, ss, n = port.find_coarse_timing(i_s, q_s, fsi_, 20)
        assert err == 0 and n == n_r and np.array_equal(ss[:n], np.asarray(ss_r)[:n]), tag
        assert np.allclose(fo[:n], np.asarray(fo_r)[:n], atol=1.0), tag
        err, ss2_r, nid2_r, ps_r, th_r, f_r = phy.find_pss(i_s, q_s, ss_r[0])
        ss_p, nid2_p, ps_p, th_p, f_p = port.find_pss(i_s, q_s, fsi_, nrb, ss[0])
        assert err == 0 and (nid2_p, ps_p, f_p) == (nid2_r, ps_r, f_r) and nid2_p == nid2 and np.array_equal(ss_p, ss2_r), tag
        assert abs(th_p - th_r) <= 1e-4 * th_r, tag
        err, ss3_r, nid1_r, fsi_r = phy.find_sss(i_s, q_s, nid2_r, ss2_r, th_r)
        rc, ss_s, nid1_p, fsi_p = port.find_sss(i_s, q_s, fsi_, nrb, nid2_p, ss_p, th_p)
        assert rc == err == 0 and nid1_p == nid1_r == nid1 and fsi_p == fsi_r and np.array_equal(ss_s, ss3_r), tag


def fs_probe():
    """oracle/_ref/libfs_probe.so: the reference's own file-scan block (fs.cc compiled unmodified) with handles on two private helpers"""
    import ctypes as C
    path = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "libfs_probe.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libfs_probe.so not built here (needs /root/reference)")
    return C.CDLL(path)


def test_freq_shift_and_capture_format_against_the_reference_block(port):
    """LTE_fdd_dl_fs_samp_buf::freq_shift (fs.cc:696-713), the reference's own object code: the port's restatement is BIT-identical for
    every sample rate, positive / negative / zero offsets and windows inside a larger buffer (sample numbering from the buffer start, float
    product (i+1)*f promoted to double by M_PI).  And copy_input_to_samp_buf (fs.cc:657-694): an int8 capture is I, Q, I, Q ... as floats,
    also when a chunk ends between an I and its Q -- the format the library's int8 loader (ofdm.cu) assumes."""
    import ctypes as C
    lib = fs_probe()
    rng = np.random.default_rng(696)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for fs in (30720000, 15360000, 7680000, 3840000, 1920000):
        for f in (0.0, 1234.5, -733.25, 7499.0, -15000.0, 3.0e-3):
            n = int(rng.integers(2000, 70000)); start = int(rng.integers(0, n // 3)); num = int(rng.integers(1, n - start))
            i0 = (rng.standard_normal(n) * 50).astype(np.float32); q0 = (rng.standard_normal(n) * 50).astype(np.float32)
            ir, qr = i0.copy(), q0.copy()
            lib.fs_probe_freq_shift(p(ir), p(qr), C.c_uint(start), C.c_uint(num), C.c_float(f), C.c_uint(fs))
            ip, qp = port.freq_shift(i0, q0, start, num, f, fs)
            assert np.array_equal(ip, ir) and np.array_equal(qp, qr), (fs, f, n, start, num)
            assert np.array_equal(ir[:start], i0[:start]) and np.array_equal(qr[start + num:], q0[start + num:])
    # large sample indices (the scan shifts up to 10 frames = 3,072,000 samples at once, fs.cc:296): the float product (i+1)*f loses integer precision there
    n = 3_100_000
    i0 = rng.standard_normal(n).astype(np.float32); q0 = rng.standard_normal(n).astype(np.float32)
    ir, qr = i0.copy(), q0.copy()
    lib.fs_probe_freq_shift(p(ir), p(qr), C.c_uint(0), C.c_uint(n), C.c_float(5321.7), C.c_uint(30720000))
    ip, qp = port.freq_shift(i0, q0, 0, n, 5321.7, 30720000)
    assert np.array_equal(ip, ir) and np.array_equal(qp, qr)
    # capture format
    items = rng.integers(-128, 128, 20001).astype(np.int8)
    for chunks in ([20001], [7, 1, 2, 9990, 10001], [1] * 9 + [19992], [4096, 4095, 4097, 7713]):
        assert sum(chunks) == items.size
        cs = np.array(chunks, np.int32)
        oi = np.zeros(items.size, np.float32); oq = np.zeros(items.size, np.float32)
        n_c = lib.fs_probe_copy_input_int8(p(items), p(cs), len(chunks), p(oi), p(oq), items.size)
        assert n_c == items.size // 2
        assert np.array_equal(oi[:n_c], items[0:2 * n_c:2].astype(np.float32)) and np.array_equal(oq[:n_c], items[1:2 * n_c:2].astype(np.float32)), chunks

"""Host logic of the product that needs no device: the tables the symbol-level and rate-unmatching kernels work through are built on
the host (openlte_b200/csrc/pdsch.cu: re_walk, rate_unmatch.cu: build_first_map) and exported for inspection.  Here they are compared
with the oracle's restatement of the reference's loops (which tests/test_oracle_port.py pins to the reference itself)."""
import ctypes as C
import os
import re

import numpy as np

import openlte_b200
from openlte_b200.lib import Alloc

HERE = os.path.dirname(os.path.abspath(__file__))


def product_re_indices(lib, n_rb, cell, n_ant, subfr, cfi, prb0, prb1):
    al = Alloc.make(1, prb0, 100, 1, prb_slot1=prb1)
    out = np.zeros(14 * 12 * len(prb0), np.int32); n = C.c_int(0)
    rc = lib.pdsch_re_indices(n_rb, cell, n_ant, subfr, cfi, C.byref(al), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n))
    assert rc == 0
    return out[:n.value]


def test_pdsch_re_index_lists_equal_the_oracle_walk(port):
    """RE walk of liblte_phy.cc:3745-3802 (CRS of 1 / 2 / 4 ports, PBCH in subframe 0, PSS/SSS in subframes 0 and 5, control symbols):
    every bandwidth, random PRB sets that differ between the slots, random cells, all subframe numbers."""
    lib = openlte_b200.load()
    rng = np.random.default_rng(31)
    n_checked = 0
    for n_rb in (6, 15, 25, 50, 75, 100):
        nsc = 12 * n_rb
        for trial in range(40):
            cell = int(rng.integers(0, 504)); n_ant = int(rng.choice([1, 2, 4])); subfr = int(rng.integers(0, 10)); cfi = int(rng.integers(1, 5))
            n_prb = int(rng.integers(1, n_rb + 1))
            prb0 = np.sort(rng.choice(n_rb, n_prb, replace=False)).astype(np.int32)
            prb1 = prb0 if trial % 3 else rng.permutation(prb0).astype(np.int32)          # distributed-style: other PRBs in slot 1
            if trial % 5 == 0:   # make sure the centre band (PBCH / PSS / SSS exclusions) is hit
                prb0 = np.arange(n_rb, dtype=np.int32); prb1 = prb0
            want = port.pdsch_re_indices(n_rb, cell, n_ant, subfr, cfi, prb0, prb1)
            got = product_re_indices(lib, n_rb, cell, n_ant, subfr, cfi, prb0, prb1)
            want = (want // 1200) * nsc + want % 1200                                   # the oracle addresses rows of 1200 subcarriers
            assert np.array_equal(got, want), (n_rb, cell, n_ant, subfr, cfi, len(prb0))
            assert got.size == np.unique(got).size and (got // nsc >= cfi).all()
            n_checked += 1
    assert n_checked == 240
    # argument checks
    al = Alloc.make(1, [0, 200], 100, 1); out = np.zeros(4000, np.int32); n = C.c_int(0)
    assert lib.pdsch_re_indices(100, 0, 1, 1, 2, C.byref(al), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)) == -1       # PRB out of range
    al = Alloc.make(1, [0, 1], 100, 1)
    assert lib.pdsch_re_indices(100, 0, 3, 1, 2, C.byref(al), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)) == -1       # 3 ports
    assert lib.pdsch_re_indices(100, 0, 1, 1, 2, C.byref(al), out.ctypes.data_as(C.c_void_p), 10, C.byref(n)) == -1 and n.value > 10   # capacity


def unmatch_through_map(first, period, e, D):
    """What rate_unmatch_kernel does with the map (one output per thread: first visit, then one every `period` e-bits, soft combining with
    the reference's sentinel rules, liblte_phy.cc:11407-11412), in numpy float32."""
    E = e.size
    out = np.full(3 * D, 10000.0, np.float32)
    k = first.astype(np.int64).copy()
    live = (k >= 0) & (k < E)
    first_visit = True
    while live.any():
        v = e[k[live]]
        w = out[live]
        if first_visit:
            w = v.copy()
        else:
            at_null = w == np.float32(10000.0)
            add = (~at_null) & (v != np.float32(10000.0))
            w = np.where(at_null, v, np.where(add, (w + v).astype(np.float32), w)).astype(np.float32)
        out[live] = w
        first_visit = False
        k = k + period
        live = live & (k < E)
    return out


def test_rate_unmatch_map_equals_the_oracle_for_every_block_size(port):
    """Gather map of every legal block size and redundancy version (sub-block interleaver, <NULL> pads, k0, circular-buffer wrap,
    liblte_phy.cc:11283-11400): applying it to punctured, exact and repeated e-sequences (with sentinel entries among them) gives exactly
    what the oracle's restatement of liblte_phy_rate_unmatch_turbo gives."""
    lib = openlte_b200.load()
    rows = re.findall(r"\{\s*(\d+),\s*(\d+),\s*(\d+)\}", open(os.path.join(HERE, "..", "openlte_b200", "csrc", "qpp_table.h")).read())
    Ks = [int(a) for a, _, _ in rows]
    assert len(Ks) == 188
    rng = np.random.default_rng(9)
    for n, K in enumerate(Ks):
        D = K + 4
        for rv in range(4):
            first = np.zeros(3 * D, np.int32); period = C.c_int(0)
            assert lib.rate_unmatch_map(K, 1, 1, 250368, 8, 0, rv, first.ctypes.data_as(C.c_void_p), C.byref(period)) == 0
            assert period.value == 3 * D                                   # one lap visits every non-<NULL> position once (N_cb = K_w here)
            assert np.array_equal(np.sort(first), np.arange(3 * D))       # ... so the map is a permutation
            if (n + rv) % 4:
                continue                                                  # the numeric comparison on every fourth (size, rv) pair
            for E in (2 * D + 7, 3 * D, 7 * D + 11):
                e = np.round(rng.standard_normal(E) * 20).astype(np.float32)
                e[rng.integers(0, E, E // 17)] = 10000.0
                want = port.rate_unmatch_turbo(e, D, rv=rv)
                assert np.array_equal(unmatch_through_map(first, period.value, e, D), want), (K, rv, E)
    # limited soft buffer (N_cb < K_w: category-1 UE with several code blocks): positions beyond N_cb are never sent
    K = 6144; D = K + 4
    first = np.zeros(3 * D, np.int32); period = C.c_int(0)
    assert lib.rate_unmatch_map(K, 3, 1, 250368, 8, 0, 0, first.ctypes.data_as(C.c_void_p), C.byref(period)) == 0
    assert period.value < 3 * D and (first == -1).sum() == 3 * D - period.value
    e = np.round(rng.standard_normal(2 * D) * 20).astype(np.float32)
    assert np.array_equal(unmatch_through_map(first, period.value, e, D), port.rate_unmatch_turbo(e, D, rv=0, C=3))
    assert lib.rate_unmatch_map(30, 1, 1, 250368, 8, 0, 0, first.ctypes.data_as(C.c_void_p), C.byref(period)) == -1
    assert lib.rate_unmatch_map(K, 1, 1, 250368, 8, 0, 4, first.ctypes.data_as(C.c_void_p), C.byref(period)) == -1


def e2e_plan(lib, n_sf, n_alloc, K, mode, starts, fs_idx=4, n_samps=None, n_sm=148):
    starts = np.ascontiguousarray(starts, np.int64)
    if n_samps is None:
        n_samps = int(starts[-1]) + (1920 << fs_idx) + 8192
    bounds = np.zeros(n_sf + 2, np.int32); ends = np.zeros(n_sf + 2, np.int64); n = C.c_int(0)
    rc = lib.e2e_plan(n_sf, n_alloc, K, mode, n_sm, fs_idx, n_samps, starts.ctypes.data_as(C.c_void_p), bounds.ctypes.data_as(C.c_void_p),
                      ends.ctypes.data_as(C.c_void_p), bounds.size, C.byref(n))
    assert rc == 0
    return bounds[:n.value + 1].copy(), ends[:n.value].copy(), n_samps


def test_host_pipeline_plan():
    """Chunk schedule and copy bounds of lteb200_dl_receive_host (capi_chain.cu: plan_chunks, chunk_copy_end) without a device: the shipped
    schedule for the bench workload is the one measured in profiles/r2_e2e_schedule.txt; chunks tile the batch; and before a chunk's kernels
    start, the cumulative copy covers everything each of its subframes reads -- with contiguous, gapped, duplicate and overlapping start
    indices (the round-1 advisor's race)."""
    lib = openlte_b200.load()
    for v in ("LTEB200_E2E_SCHEDULE", "LTEB200_E2E_CHUNK"):
        os.environ.pop(v, None)
    n_sf_samps = 30720
    starts = np.arange(4000, dtype=np.int64) * n_sf_samps
    # the bench workload: 4000 subframes x 4 transport blocks of K = 6144 -> wave = 148 SMs x 2 blocks / 4 = 74 subframes
    b, ends, n_samps = e2e_plan(lib, 4000, 4, 6144, 1, starts)
    assert len(b) - 1 == 10 and b[1] == 148                                    # "10 chunks, first 148 subframes" in the measured trace
    sizes = np.diff(b)
    assert (sizes[:-2] % 74 == 0).all() and sizes.max() <= 12 * 74 and sizes[-1] == 3 * 74 and (np.diff(sizes[:-2]) >= 0).all()
    # the reference-exact decoder gets few large chunks: 6, 14, 20 waves (profiles/r2_bench_mode_a.json)
    b0, _, _ = e2e_plan(lib, 4000, 4, 6144, 0, starts)
    assert np.diff(b0)[:3].tolist() == [444, 1036, 1480] and len(b0) - 1 <= 5
    # small batches: one chunk
    b1, e1, ns1 = e2e_plan(lib, 100, 4, 6144, 1, starts[:100])
    assert b1.tolist() == [0, 100] and e1.tolist() == [ns1]
    rng = np.random.default_rng(12)
    look = {4: 2048 + 160 + 2048 + 144}                                       # two look-ahead symbols at 30.72 Msps (liblte_phy.cc:5946-5957)
    for trial in range(60):
        n_sf = int(rng.integers(1, 9000)); n_alloc = int(rng.integers(1, 9)); K = int(rng.choice([40, 1024, 2048, 3072, 6144])); mode = int(rng.integers(0, 2))
        kind = trial % 4
        if kind == 0:
            st = np.arange(n_sf, dtype=np.int64) * n_sf_samps + int(rng.integers(0, 5000))
        elif kind == 1:      # gaps
            st = np.cumsum(n_sf_samps + rng.integers(0, 3 * n_sf_samps, n_sf)).astype(np.int64)
        elif kind == 2:      # duplicates and overlaps
            st = np.cumsum(rng.choice([0, 17, 1000, n_sf_samps], n_sf)).astype(np.int64)
        else:                # long runs of one index, then a jump
            st = np.repeat(np.arange((n_sf + 99) // 100, dtype=np.int64) * 5 * n_sf_samps, 100)[:n_sf]
        n_samps = int(st[-1]) + n_sf_samps + int(rng.choice([0, 100, 4400, 9000]))      # streams that end inside the last look-ahead too
        b, ends, _ = e2e_plan(lib, n_sf, n_alloc, K, mode, st, n_samps=n_samps)
        assert b[0] == 0 and b[-1] == n_sf and (np.diff(b) > 0).all(), (trial, b)
        assert (np.diff(ends) >= 0).all() and ends[-1] == n_samps
        for c in range(len(b) - 1):
            need = min(n_samps, int(st[b[c]:b[c + 1]].max()) + n_sf_samps + look[4])
            assert ends[c] >= need, (trial, kind, c, ends[c], need)
    # development overrides
    os.environ["LTEB200_E2E_CHUNK"] = "64"
    try:
        b, _, _ = e2e_plan(lib, 1000, 4, 6144, 1, starts[:1000])
        assert np.diff(b).tolist() == [64] * 15 + [40]
    finally:
        os.environ.pop("LTEB200_E2E_CHUNK", None)
    os.environ["LTEB200_E2E_SCHEDULE"] = "592,1480,1480"
    try:
        b, _, _ = e2e_plan(lib, 4000, 4, 6144, 0, starts)
        assert np.diff(b).tolist() == [592, 1480, 1480, 448]
    finally:
        os.environ.pop("LTEB200_E2E_SCHEDULE", None)
    # refusals: decreasing start indices, no room for the result
    bad = starts[:10].copy(); bad[5] = 0
    bounds = np.zeros(12, np.int32); ends = np.zeros(12, np.int64); n = C.c_int(0)
    assert lib.e2e_plan(10, 4, 6144, 1, 148, 4, 10 ** 7, bad.ctypes.data_as(C.c_void_p), bounds.ctypes.data_as(C.c_void_p), ends.ctypes.data_as(C.c_void_p), 12, C.byref(n)) == -1
    assert lib.e2e_plan(4000, 4, 6144, 1, 148, 4, 4000 * 30720 + 8192, starts.ctypes.data_as(C.c_void_p), bounds.ctypes.data_as(C.c_void_p), ends.ctypes.data_as(C.c_void_p), 5, C.byref(n)) == -1 and n.value == 10


def test_dci_1a_transport_block_sizes(port):
    """lteb200_dci_1a_tbs (what the device transmitter announces in a DCI 1A) against the oracle's copy of the reference's table
    (36.213 table 7.1.7.2.1-1 column 3 as liblte_phy.cc:13273-13380 indexes it), which tests/test_control_oracle.py pins to the reference."""
    lib = openlte_b200.load()
    assert [lib.dci_1a_tbs(m) for m in range(27)] == [port.tbs_1a(m, 3) for m in range(27)]
    assert lib.dci_1a_tbs(27) == -1 and lib.dci_1a_tbs(-1) == -1


def test_host_functions_of_the_drop_in_equal_the_reference(refo, tmp_path):
    """liblte_phy_code_block_segmentation / _desegmentation of the drop-in (host code, openlte_b200/csrc/dropin.cu) against the reference's own
    functions (liblte_phy.cc:9753-9987) on random transport blocks from one to thirteen code blocks: block count, filler count, per-block sizes
    (with the reference's "+24" quirk of the multi-block branch), every byte of the segmented blocks -- including the CRC24B of the
    first block, which the reference computes over filler bytes of value 100 -- and the desegmented bits.  Also liblte_phy_update_n_rb_dl
    (:2592-2652) for every sample rate x bandwidth pair, legal or not."""
    import subprocess
    import pytest
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    root = os.path.join(HERE, "..")
    so = str(tmp_path / "seg_shim.so")
    subprocess.run(["g++", "-shared", "-fPIC", "-DLIBLTE_PHY_B200_STANDALONE", "-I", os.path.join(root, "include"), os.path.join(HERE, "cpp", "seg_shim.cc"),
                    "-L", os.path.join(root, "openlte_b200"), "-lopenlte_b200", "-Wl,-rpath," + os.path.abspath(os.path.join(root, "openlte_b200")), "-o", so],
                   check=True, capture_output=True)
    sh = C.CDLL(so); rl = refo.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    CMAX = 6144 + 64

    def run(seg, deseg, b):
        Cn = C.c_uint(0); F = C.c_uint(0)
        c = np.zeros(16 * CMAX, np.uint8); ncb = np.zeros(16, np.uint32)
        seg(p(b), C.c_uint(b.size), C.byref(Cn), C.byref(F), p(c), C.c_uint(CMAX), p(ncb))
        out = np.zeros(b.size + 64, np.uint8)
        deseg(p(c), p(ncb), C.c_uint(CMAX), C.c_uint(b.size - 24), p(out), C.c_uint(b.size))
        return Cn.value, F.value, ncb, c, out

    rng = np.random.default_rng(24)
    sizes = [40, 64, 168, 1000, 6144, 6145, 6200, 12000, 12288, 20000, 30000, 45000, 61664, 75376 + 24] + [int(x) for x in rng.integers(40, 80000, 40)]
    n_multi = 0
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)     # the reference prints "CB CRC INCORRECT" (its own H7 bug)
    try:
        for nb in sizes:
            b = rng.integers(0, 2, nb).astype(np.uint8)
            r = run(rl.ref_code_block_segmentation, rl.ref_code_block_desegmentation, b)
            d = run(sh.shim_code_block_segmentation, sh.shim_code_block_desegmentation, b)
            assert (r[0], r[1]) == (d[0], d[1]) and np.array_equal(r[2], d[2]), (nb, r[:2], d[:2])
            assert np.array_equal(r[3], d[3]), nb
            assert np.array_equal(r[4], d[4]), nb
            if r[0] == 1:
                assert np.array_equal(d[4][:nb], b)                     # single block: the round trip is the identity
            n_multi += r[0] > 1
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)
    assert n_multi >= 30
    # liblte_phy_update_n_rb_dl (liblte_phy.cc:2592-2652): legal (sample rate, bandwidth) pairs, FFT size and padding, error code
    for n_symb in (128, 256, 512, 1024, 2048, 1536, 0):
        for n_rb in (0, 6, 10, 15, 25, 50, 75, 100, 110):
            a = np.zeros(4, np.uint32); b = np.zeros(4, np.uint32)
            rl.ref_update_n_rb_dl_probe(C.c_uint(n_symb), C.c_uint(n_rb), p(a)); sh.shim_update_n_rb_dl_probe(C.c_uint(n_symb), C.c_uint(n_rb), p(b))
            assert np.array_equal(a, b), (n_symb, n_rb, a, b)


def test_transmit_side_channel_coding_equals_the_reference(refo, port):
    """Host code of the device transmitter against the reference's own functions on random inputs: DCI format 1A packing + channel coding to 288
    bits (dci_1a_pack liblte_phy.cc:13153-13250, dci_channel_encode :12882-12942) for every bandwidth, allocation size / position, MCS,
    redundancy version and the SI / P / RA RNTIs, with the transport block size the packer derives; and the PBCH coding chain (bch_channel_encode
    :12510-12571 + the scrambling of :3894-3899) for 1 / 2 / 4 ports."""
    import pytest
    if refo is None:
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    lib = openlte_b200.load(); rl = refo.lib(); phy = refo.RefPhy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(1353)
    n = 0
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
    try:
        for n_rb in (6, 15, 25, 50, 75, 100):
            for trial in range(60):
                n_prb = int(rng.integers(1, n_rb + 1)); prb0 = int(rng.integers(0, n_rb - n_prb + 1))
                mcs = int(rng.integers(0, 27)); rv = int(rng.integers(0, 4))
                rnti = int(rng.choice([0xFFFF, 0xFFFE, int(rng.integers(1, 0x3D))]))
                al = Alloc.make(1, list(range(prb0, prb0 + n_prb)), 0, rnti, rv_idx=rv)
                got = np.zeros(288, np.uint8); want = np.zeros(288, np.uint8); tbs_ref = C.c_int(0)
                tbs = lib.dci_1a_encode(C.byref(al), mcs, n_rb, p(got))
                rl.ref_dci_1a_encode(phy.h, rnti, mcs, n_prb, prb0, rv, n_rb, p(want), C.byref(tbs_ref))
                assert np.array_equal(got, want), (n_rb, n_prb, prb0, mcs, rv, hex(rnti))
                assert tbs == tbs_ref.value == lib.dci_1a_tbs(mcs)
                n += 1
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)
    assert n == 360
    # refusals: non-contiguous PRBs, MCS beyond the table
    al = Alloc.make(1, [3, 5], 0, 0xFFFF); out = np.zeros(288, np.uint8)
    assert lib.dci_1a_encode(C.byref(al), 3, 25, p(out)) == -1
    al = Alloc.make(1, [3, 4], 0, 0xFFFF)
    assert lib.dci_1a_encode(C.byref(al), 27, 25, p(out)) == -1
    # PBCH
    for trial in range(30):
        mib = rng.integers(0, 2, 24).astype(np.uint8); n_ant = (1, 2, 4)[trial % 3]; cell = int(rng.integers(0, 504))
        got = np.zeros(1920, np.uint8); coded = np.zeros(1920, np.uint8)
        assert lib.bch_encode(p(mib), n_ant, cell, p(got)) == 0
        assert rl.ref_bch_encode_bits(phy.h, p(mib), n_ant, p(coded)) == 1920
        assert np.array_equal(got, coded ^ port.prs_c(cell, 1920)), (n_ant, cell)
    assert lib.bch_encode(p(mib), 3, 0, p(got)) == -1

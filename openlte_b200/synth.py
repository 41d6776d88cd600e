"""Synthetic LTE downlink signal generator (numpy, vectorised over subframes) -- DATA SYNTHESIS ONLY.

Builds the I/Q the receive chain is benchmarked and tested on, following the transmit side of the reference so that
both the reference and this library decode it: CRS mapping (liblte_phy_map_crs, liblte_phy.cc:5144-5256), DL-SCH
encoding for single-code-block transport blocks (dlsch_channel_encode :12664-12753: CRC24A, turbo_encode :10541-10586
with the uint32-wrapping interleaver :10982, liblte_phy_rate_match_turbo :11081-11236), scrambling (:3579-3590),
modulation_mapper (:8704-9490, 36.211 7.1), the RE walk of liblte_phy_pdsch_channel_encode (:3618-3672) and
symbols_to_samples_dl (:8484-8531: unnormalised IFFT + cyclic prefix).  Not on the product path, never timed;
tests/test_synth.py checks it against the reference's own transmit functions.
"""
import numpy as np

QM = {0: 1, 1: 2, 2: 4, 3: 6}
_P32 = np.array([0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30, 1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31])


def _qpp_rows():
    import os, re
    txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "qpp_table.h")).read()
    return {int(a): (int(b), int(c)) for a, b, c in re.findall(r"\{\s*(\d+),\s*(\d+),\s*(\d+)\}", txt)}


_QPP = None


def qpp_pi(K):
    """pi(i) with the reference's uint32 wrap-around."""
    global _QPP
    if _QPP is None:
        _QPP = _qpp_rows()
    f1, f2 = _QPP[K]
    i = np.arange(K, dtype=np.uint64)
    return (((f1 * i + f2 * i * i) & 0xFFFFFFFF) % K).astype(np.int64)


def seg_K(tbs):
    B = tbs + 24
    for lo, hi, st in ((40, 512, 8), (528, 1024, 16), (1056, 2048, 32), (2112, 6144, 64)):
        for k in range(lo, hi + 1, st):
            if k >= B:
                return k
    raise ValueError("tbs needs more than one code block")


def prs_c(c_init, n):
    """Gold sequence c(n) of 36.211 7.2 (generate_prs_c, liblte_phy.cc:9669-9704), generated 28 elements at a time."""
    total = 1600 + n
    x1 = np.zeros(total + 96, np.uint8); x2 = np.zeros(total + 96, np.uint8)
    x1[0] = 1
    x2[:31] = [(c_init >> i) & 1 for i in range(31)]
    for base in range(0, total, 28):
        sl = slice(base, base + 28)
        x1[base + 31:base + 59] = x1[base + 3:base + 31] ^ x1[sl]
        x2[base + 31:base + 59] = x2[base + 3:base + 31] ^ x2[base + 2:base + 30] ^ x2[base + 1:base + 29] ^ x2[sl]
    return (x1[1600:1600 + n] ^ x2[1600:1600 + n]).astype(np.uint8)


def crc_bits(bits, poly=0x1864CFB, n_p=24):
    """bits: [B, n] -> parity [B, n_p] (calc_crc, liblte_phy.cc:9713-9743), vectorised across B."""
    bits = np.atleast_2d(bits).astype(np.uint32)
    B, n = bits.shape
    rem = np.zeros(B, np.uint32)
    top = np.uint32(1 << n_p)
    for i in range(n + n_p):
        rem = (rem << np.uint32(1)) | (bits[:, i] if i < n else np.uint32(0))
        rem = np.where(rem & top, rem ^ np.uint32(poly), rem)
    return ((rem[:, None] >> np.arange(n_p - 1, -1, -1, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)


def _rsc(x):
    """x: [B, K] bits -> (z [B, K+4], fb [B, K+4]); feedback 1+D^2+D^3, feed-forward 1+D+D^3, 4 termination steps."""
    B, K = x.shape
    z = np.zeros((B, K + 4), np.uint8); fb = np.zeros((B, K + 4), np.uint8)
    s1 = np.zeros(B, np.uint8); s2 = np.zeros(B, np.uint8); s3 = np.zeros(B, np.uint8)
    for i in range(K + 4):
        f = s2 ^ s3
        a = (x[:, i] ^ f) if i < K else np.zeros(B, np.uint8)
        fb[:, i] = f
        z[:, i] = a ^ s1 ^ s3
        s3, s2, s1 = s2, s1, a
    return z, fb


def turbo_encode(c):
    """c: [B, K] -> d: [B, 3, K+4] branch-major (turbo_encode, liblte_phy.cc:10541-10586)."""
    c = np.atleast_2d(c).astype(np.uint8)
    B, K = c.shape
    pi = qpp_pi(K)
    z, fb = _rsc(c)
    zp, xp = _rsc(c[:, pi])
    d = np.zeros((B, 3, K + 4), np.uint8)
    d[:, 0, :K] = c; d[:, 1, :K] = z[:, :K]; d[:, 2, :K] = zp[:, :K]
    d[:, 0, K:] = np.stack([fb[:, K], z[:, K + 1], xp[:, K], zp[:, K + 1]], 1)
    d[:, 1, K:] = np.stack([z[:, K], fb[:, K + 2], zp[:, K], xp[:, K + 2]], 1)
    d[:, 2, K:] = np.stack([fb[:, K + 1], z[:, K + 2], xp[:, K + 1], zp[:, K + 2]], 1)
    return d


def rate_match_map(D, E, rv=0, C=1, tx_mode=1, N_soft=250368, M_dl_harq=8):
    """index into the branch-major d (x*D + t) of every transmitted bit e[k] (36.212 5.1.4.1)."""
    R = (D + 31) // 32; Kpi = 32 * R; ND = Kpi - D; Kw = 3 * Kpi
    n = np.arange(Kpi); col = n // R; row = n % R
    t01 = 32 * row + _P32[col]; t2 = (_P32[col] + 32 * row + 1) % Kpi
    sx = np.zeros(Kw, np.int64); st = np.zeros(Kw, np.int64)
    sx[:Kpi] = 0; st[:Kpi] = t01
    sx[Kpi::2] = 1; st[Kpi::2] = t01
    sx[Kpi + 1::2] = 2; st[Kpi + 1::2] = t2
    K_mimo = 2 if tx_mode in (3, 4, 8, 9) else 1
    N_ir = N_soft // (K_mimo * min(M_dl_harq, 8))
    Ncb = min(N_ir // C, Kw)
    k0 = int(R * (2 * int(np.ceil(Ncb / (8 * R))) * rv + 2))
    pos = (k0 + np.arange(Ncb)) % Ncb
    pos = pos[st[pos] >= ND]
    reps = -(-E // pos.size)
    pos = np.tile(pos, reps)[:E]
    return sx[pos] * D + (st[pos] - ND)


def mod_map(bits, mod):
    """bits [..., n*Qm] -> complex64 symbols (36.211 7.1.1-7.1.4)."""
    b = 1.0 - 2.0 * bits.astype(np.float32)
    if mod == 0:
        s = b / np.float32(np.sqrt(2.0)); return (s + 1j * s).astype(np.complex64)
    if mod == 1:
        return ((b[..., 0::2] + 1j * b[..., 1::2]) * np.float32(1 / np.sqrt(2.0))).astype(np.complex64)
    if mod == 2:
        i = b[..., 0::4] * (2 - b[..., 2::4]); q = b[..., 1::4] * (2 - b[..., 3::4])
        return ((i + 1j * q) * np.float32(1 / np.sqrt(10.0))).astype(np.complex64)
    i = b[..., 0::6] * (4 - b[..., 2::6] * (2 - b[..., 4::6])); q = b[..., 1::6] * (4 - b[..., 3::6] * (2 - b[..., 5::6]))
    return ((i + 1j * q) * np.float32(1 / np.sqrt(42.0))).astype(np.complex64)


def re_indices(N_rb_dl, N_id_cell, subfr, cfi, prb):
    """(symbol, subcarrier) walk of liblte_phy.cc:3618-3672 for one antenna port."""
    centre = {6: (0, 71), 15: (54, 125), 25: (114, 185), 50: (264, 335), 75: (414, 485)}.get(N_rb_dl, (564, 635))
    out = []
    for L in range(cfi, 14):
        for i in prb:
            for j in range(12):
                sc = 12 * i + j; l7 = L % 7
                if l7 == 0 and N_id_cell % 6 == j % 6: continue
                if l7 == 4 and (N_id_cell + 3) % 6 == j % 6: continue
                c = centre[0] <= sc <= centre[1]
                if subfr == 0 and c and 7 <= L <= 10: continue
                if subfr in (0, 5) and c and L in (5, 6): continue
                out.append((L, sc))
    return np.array(out, np.int64)


def crs_grid(N_rb_dl, N_id_cell, subfr):
    """[14, 12*N_rb] complex64 with the port-0 CRS of one subframe, zeros elsewhere."""
    nsc = 12 * N_rb_dl
    g = np.zeros((14, nsc), np.complex64)
    r2 = np.float32(1 / np.sqrt(2.0))
    for sym, slot, l, v in ((0, 0, 0, 0), (4, 0, 4, 3), (7, 1, 0, 0), (11, 1, 4, 3)):
        ns = 2 * subfr + slot
        c = prs_c(1024 * (7 * (ns + 1) + l + 1) * (2 * N_id_cell + 1) + 2 * N_id_cell + 1, 440).astype(np.float32)
        val = (r2 * (1 - 2 * c[0::2]) + 1j * r2 * (1 - 2 * c[1::2])).astype(np.complex64)
        j = np.arange(2 * N_rb_dl)
        g[sym, 6 * j + (v + N_id_cell % 6) % 6] = val[j + 110 - N_rb_dl]
    return g


def ofdm_modulate(grid, fs_idx):
    """grid [B, 14, nsc] -> samples [B, n_sf] complex64 (symbols_to_samples_dl: unnormalised IFFT, CP prepended)."""
    B, _, nsc = grid.shape
    N = 128 << fs_idx; cp0 = 10 << fs_idx; cpe = 9 << fs_idx; h = nsc // 2
    X = np.zeros((B, 14, N), np.complex64)
    X[:, :, 1:h + 1] = grid[:, :, h:]
    X[:, :, N - h:] = grid[:, :, :h]
    x = (np.fft.ifft(X, axis=-1) * N).astype(np.complex64)
    out = np.zeros((B, 1920 << fs_idx), np.complex64)
    o = 0
    for s in range(14):
        cp = cp0 if s % 7 == 0 else cpe
        out[:, o:o + cp] = x[:, s, N - cp:]
        out[:, o + cp:o + cp + N] = x[:, s]
        o += cp + N
    return out


def make_subframes(n_sf, allocs, N_id_cell=0, N_rb_dl=100, fs_idx=4, cfi=2, subfr_nums=None, seed=0, gain=None, noise=0.0, tb_bits=None,
                   first_subfr=1):
    """n_sf independent subframes, each carrying every allocation in `allocs` (dict mod, prb, tbs, rnti[, rv]).

    Returns dict(samples [n_sf, n] complex64 -- each row is one subframe followed by nothing (callers lay rows out in a
    stream and append a CRS-only tail), tb [n_sf][n_alloc] bit arrays, subfr_nums)."""
    rng = np.random.default_rng(seed)
    if subfr_nums is None:
        # consecutive numbering: in a contiguous stream the estimator reads symbols 14-15 from the NEXT subframe and
        # expects the CRS of slot (2*n+2) mod 20 there (liblte_phy.cc:5966-5967)
        subfr_nums = ((first_subfr + np.arange(n_sf)) % 10).astype(np.int32)
    nsc = 12 * N_rb_dl
    grid = np.zeros((n_sf, 14, nsc), np.complex64)
    crs_by_sn = {}
    for sn in np.unique(subfr_nums):
        crs_by_sn[int(sn)] = crs_grid(N_rb_dl, N_id_cell, int(sn))
    for b in range(n_sf):
        grid[b] = crs_by_sn[int(subfr_nums[b])]
    tbs_out = []
    for ai, a in enumerate(allocs):
        tbs = a["tbs"]; K = seg_K(tbs); D = K + 4; mod = a["mod"]; Qm = QM[mod]
        assert K == tbs + 24, "filler bits are not handled by the reference encoder; use a TBS with TBS+24 in the K table"
        tb = rng.integers(0, 2, (n_sf, tbs)).astype(np.uint8) if tb_bits is None else np.asarray(tb_bits[ai], np.uint8).reshape(n_sf, tbs)
        tbs_out.append(tb)
        cw = np.concatenate([tb, crc_bits(tb)], axis=1)
        d = turbo_encode(cw).reshape(n_sf, 3 * D)
        for sn in np.unique(subfr_nums):
            sel = np.nonzero(subfr_nums == sn)[0]
            re = re_indices(N_rb_dl, N_id_cell, int(sn), cfi, a["prb"])
            G = re.shape[0] * Qm
            e = d[sel][:, rate_match_map(D, G, rv=a.get("rv", 0))]
            c = prs_c((a["rnti"] << 14) | (int(sn) << 9) | N_id_cell, G)
            sym = mod_map(e ^ c[None, :], mod)
            grid[sel[:, None], re[None, :, 0], re[None, :, 1]] = sym
    samples = ofdm_modulate(grid, fs_idx)
    if gain is not None:
        samples = samples * np.asarray(gain, np.complex64).reshape(-1, 1)
    if noise > 0:
        samples = samples + noise * (rng.standard_normal(samples.shape) + 1j * rng.standard_normal(samples.shape)).astype(np.complex64)
    return dict(samples=samples.astype(np.complex64), tb=tbs_out, subfr_nums=subfr_nums, grid=grid)


def crs_only_tail(N_rb_dl, N_id_cell, subfr, fs_idx):
    """First two OFDM symbols' worth (a whole CRS-only subframe is returned) for the estimator's look-ahead."""
    return ofdm_modulate(crs_grid(N_rb_dl, N_id_cell, subfr)[None], fs_idx)[0]

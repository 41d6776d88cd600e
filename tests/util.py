"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def noisy_codeword(port, K, sigma, rng, sentinel_every=0):
    """Random block -> reference-compatible turbo code word -> BPSK + AWGN in the decoder's d[3i+x] layout."""
    c = rng.integers(0, 2, K).astype(np.uint8)
    d = port.turbo_encode(c)                       # branch-major
    D = K + 4
    x = (1.0 - 2.0 * d.astype(np.float32)).reshape(3, D).T.reshape(-1).copy()
    if sigma > 0:
        x = (x + sigma * rng.standard_normal(x.size)).astype(np.float32)
    if sentinel_every:
        x[1::sentinel_every] = 10000.0
    return c, x.astype(np.float32)


def dptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import ctypes
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------------------------
# Synthetic downlink signal through the REFERENCE's own transmit functions: CRS +
# PDSCH allocations -> IFFT, followed by a CRS-only subframe (the estimator reads symbols 14-15).
def ref_loopback(refo, subfr, allocs, N_id_cell=0, cfi=2, N_rb_dl=100, fs=30.72, seed=0, gain=1.0 + 0j, noise=0.0,
                 frame_start=0, n_extra=1, N_ant=1, ant_gains=None, bch=None, pdcch=None):
    """allocs: list of dict(mod, prb, tbs, rnti).  bch: dict(bits[24], sfn) -> PBCH in the first subframe (use subfr=0);
    pdcch: dict(cfi, allocs=[dict(rnti, mcs, N_prb, prb0)]) -> PCFICH + DCI 1A in the first subframe.
    Returns dict with i/q sample arrays (float32), tb bits, phy handle."""
    rng = np.random.default_rng(seed)
    phy = refo.RefPhy(fs=fs, N_id_cell=N_id_cell, N_ant=N_ant, N_rb_dl=N_rb_dl)
    if ant_gains is None:
        ant_gains = [1.0, 0.8 * np.exp(0.7j), 0.6 * np.exp(-1.9j), 0.9 * np.exp(2.5j)][:N_ant]
    n_sf = {30.72: 30720, 15.36: 15360, 7.68: 7680, 3.84: 3840, 1.92: 1920}[fs]
    total = frame_start + (subfr + 1 + n_extra) * n_sf + 4096
    i_s = np.zeros(total, np.float32); q_s = np.zeros(total, np.float32)
    tbs_bits = []
    for k in range(1 + n_extra):
        sf = refo.Subframe(); sf.clear(); sf.set_num((subfr + k) % 10)
        phy.map_crs(sf, N_id_cell, N_ant)
        if k == 0:
            for a in allocs:
                tb = rng.integers(0, 2, a["tbs"]).astype(np.uint8)
                tbs_bits.append(tb)
                phy.pdsch_encode(sf, cfi, N_id_cell, N_ant, a["mod"], a["prb"], a["tbs"], a["rnti"], tb, rv=a.get("rv", 0))
            if bch is not None:
                assert phy.bch_encode(sf, bch["bits"], N_id_cell, N_ant, bch["sfn"]) == 0
            if pdcch is not None:
                assert phy.pdcch_encode(sf, N_id_cell, N_ant, pdcch["cfi"], pdcch["allocs"], phich_res=pdcch.get("phich_res", 1.0)) == 0
        o = frame_start + (subfr + k) * n_sf
        for ant in range(N_ant):          # flat per-antenna channel, signals add at the receiver
            i_sf, q_sf = phy.create_dl_subframe(sf, ant, n_sf)
            za = (i_sf + 1j * q_sf) * ant_gains[ant]
            i_s[o:o + n_sf] += za.real.astype(np.float32); q_s[o:o + n_sf] += za.imag.astype(np.float32)
    z = (i_s + 1j * q_s) * gain
    if noise > 0:
        z = z + noise * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))
    return dict(i=np.ascontiguousarray(z.real.astype(np.float32)), q=np.ascontiguousarray(z.imag.astype(np.float32)),
                tb=tbs_bits, phy=phy, n_sf=n_sf, frame_start=frame_start)

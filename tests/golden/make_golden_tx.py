"""Generates tests/golden/tx_frames_ref.npz with the REFERENCE's own transmit functions (oracle/_ref): resource grids of subframes that carry
everything LTE_fdd_dl_file_gen puts on the air -- CRS of every port (liblte_phy_map_crs), PSS / SSS (liblte_phy_map_pss / _sss), PBCH
(liblte_phy_bch_channel_encode), PCFICH + PDCCH with DCI 1A (liblte_phy_pdcch_channel_encode) and PDSCH (liblte_phy_pdsch_channel_encode) -- for one,
two and four antenna ports.  The device-side generator (lteb200_dl_transmit_frames_batch) must reproduce these grids bit for bit
(tests/test_tx_frames_gpu.py).  Run where /root/reference exists:  python tests/golden/make_golden_tx.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as refo

FS = {100: 30.72, 50: 15.36, 25: 7.68}
TBS3 = [56, 88, 144, 176, 208, 224, 256, 328, 392, 456, 504, 584, 680, 744, 840, 904, 968, 1064, 1160, 1288, 1384, 1480, 1608, 1736, 1800, 1864, 2216]


def case(name, n_rb, cell, n_ant, cfi, sfn0, subframes, allocs, pdcch, seed):
    """subframes: list of (frame offset, subframe number); allocs: dict(rnti, mcs or tbs, mod, prb0, N_prb)"""
    rng = np.random.default_rng(seed)
    phy = refo.RefPhy(fs=FS[n_rb], N_id_cell=cell, N_ant=n_ant, N_rb_dl=n_rb)
    n_frames = max(f for f, _ in subframes) + 1
    mib = rng.integers(0, 2, (n_frames, 24)).astype(np.uint8)
    out = {f"{name}_mib": mib, f"{name}_meta": np.array([n_rb, cell, n_ant, cfi, sfn0, int(pdcch)], np.int32),
           f"{name}_subframes": np.array(subframes, np.int32)}
    grids, tbs_all = [], []
    for f, sn in subframes:
        sf = refo.Subframe(); sf.clear(); sf.set_num(sn)
        phy.map_crs(sf, cell, n_ant)
        if sn in (0, 5):
            phy.map_pss(sf, cell % 3, n_ant)
            phy.map_sss(sf, cell // 3, cell % 3, n_ant)
        if sn == 0:
            assert phy.bch_encode(sf, mib[f], cell, n_ant, sfn0 + f) == 0
        if pdcch:
            assert phy.pdcch_encode(sf, cell, n_ant, cfi, [dict(rnti=a["rnti"], mcs=a["mcs"], rv_idx=0, N_prb=a["N_prb"], prb0=a["prb0"]) for a in allocs]) == 0
        tbs_sf = []
        for a in allocs:
            tb = rng.integers(0, 2, a["tbs"]).astype(np.uint8)
            tbs_sf.append(np.pad(tb, (0, 2216 - a["tbs"])))
            n_sym = cfi + (1 if n_rb <= 10 else 0)
            assert phy.pdsch_encode(sf, n_sym, cell, n_ant, a["mod"], list(range(a["prb0"], a["prb0"] + a["N_prb"])), a["tbs"], a["rnti"], tb, rv=0,
                                    tx_mode=1 if n_ant == 1 else 2) == 0
        nsc = 12 * n_rb
        g = (sf.tx_symb_re[:n_ant, :14, :nsc] + 1j * sf.tx_symb_im[:n_ant, :14, :nsc]).astype(np.complex64)
        grids.append(g); tbs_all.append(np.stack(tbs_sf))
    out[f"{name}_grid"] = np.stack(grids)            # [n_sf][n_ant][14][nsc]
    out[f"{name}_tb"] = np.stack(tbs_all)            # [n_sf][n_alloc][2216]
    out[f"{name}_allocs"] = np.array([[a["rnti"], a["mcs"], a["tbs"], a["mod"], a["prb0"], a["N_prb"]] for a in allocs], np.int32)
    return out


def main():
    d = {}
    # A: one port, 20 MHz, everything incl. PDCCH; frames sfn 5..6 (sfn mod 4 = 1, 2): subframes 0, 5 of frame 0 and 0 of frame 1
    a_allocs = [dict(rnti=0xFFFF, mcs=9, tbs=TBS3[9], mod=1, prb0=10, N_prb=4), dict(rnti=0x0002, mcs=20, tbs=TBS3[20], mod=1, prb0=40, N_prb=9),
                dict(rnti=0xFFFE, mcs=2, tbs=TBS3[2], mod=1, prb0=60, N_prb=3)]
    d.update(case("a", 100, 17, 1, 2, 5, [(0, 0), (0, 3), (0, 5), (1, 0)], a_allocs, True, 1))
    # B: two ports, 10 MHz, sync + PBCH + PDSCH (16QAM tbs 2024 -> K 2048; 64QAM), no PDCCH (undefined in the reference beyond one port)
    b_allocs = [dict(rnti=0x123, mcs=0, tbs=2024, mod=2, prb0=5, N_prb=20), dict(rnti=0x77, mcs=0, tbs=1000, mod=3, prb0=30, N_prb=6)]
    d.update(case("b", 50, 301, 2, 3, 3, [(0, 0), (0, 7), (1, 0), (1, 5)], b_allocs, False, 2))
    # C: four ports, 5 MHz
    c_allocs = [dict(rnti=0x45, mcs=0, tbs=488, mod=1, prb0=2, N_prb=12)]
    d.update(case("c", 25, 400, 4, 2, 0, [(0, 0), (0, 1), (0, 5)], c_allocs, False, 3))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tx_frames_ref.npz"), **d)
    print({k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()

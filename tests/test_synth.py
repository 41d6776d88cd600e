"""The synthetic transmit chain used by bench.py and smoke() (openlte_b200/synth.py) against the reference's own
transmitter (CPU; needs oracle/_ref) and against the oracle port's pieces (always)."""
import numpy as np
import pytest

from openlte_b200 import synth


def test_pieces_against_port(port):
    rng = np.random.default_rng(0)
    for ci in (0, 1, (0xFFFF << 14) | (5 << 9) | 3):
        assert np.array_equal(synth.prs_c(ci, 3000), port.prs_c(ci, 3000))
    b = rng.integers(0, 2, (2, 6120)).astype(np.uint8)
    for i in range(2):
        assert int("".join(map(str, synth.crc_bits(b[i:i + 1])[0])), 2) == port.crc(b[i])
    c = rng.integers(0, 2, (2, 6144)).astype(np.uint8)
    d = synth.turbo_encode(c)
    for i in range(2):
        assert np.array_equal(d[i].reshape(-1), port.turbo_encode(c[i]))
    for K, E, rv in ((6144, 20700, 0), (1024, 3000, 2), (40, 200, 3), (6144, 6624, 0), (512, 700, 1)):     # the last two puncture
        D = K + 4
        dd = rng.integers(0, 2, 3 * D).astype(np.uint8)
        assert np.array_equal(dd[synth.rate_match_map(D, E, rv)], port.rate_match_turbo(dd, E, rv=rv))


def test_grid_and_samples_against_reference_transmitter(refo):
    if refo is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(1)
    for cell, sn, mod, prb, tbs in ((0, 1, 3, list(range(12)), 2216), (77, 5, 1, list(range(40, 76)), 3240), (150, 0, 2, list(range(41, 59)), 3112)):
        phy = refo.RefPhy(N_id_cell=cell)
        sf = refo.Subframe(); sf.clear(); sf.set_num(sn)
        tb = rng.integers(0, 2, tbs).astype(np.uint8)
        phy.map_crs(sf, cell, 1)
        phy.pdsch_encode(sf, 2, cell, 1, mod, prb, tbs, 0x4321, tb)
        want = (sf.tx_symb_re[0, :14] + 1j * sf.tx_symb_im[0, :14]).astype(np.complex64)
        r = synth.make_subframes(1, [dict(mod=mod, prb=prb, tbs=tbs, rnti=0x4321)], N_id_cell=cell, subfr_nums=np.array([sn]), tb_bits=[tb[None]])
        assert np.array_equal(r["grid"][0], want)                       # bit-identical resource grid
        i_s, q_s = phy.create_dl_subframe(sf, 0, 30720)
        z = i_s + 1j * q_s
        assert np.abs(r["samples"][0] - z).max() <= 1e-6 * np.abs(z).max()   # numpy's double IFFT vs the reference's float IFFT

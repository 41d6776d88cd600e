"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle_port.so (the CPU restatement in oracle/port/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_port.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        if hasattr(_lib, "op_crc"):
            _lib.op_crc.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_qpp_spec(on):
    """Interleaver of every port function: False = the reference's uint32-wrapping one (default), True = 3GPP TS 36.212's."""
    lib().op_set_qpp_spec(1 if on else 0)


def qpp_params(K):
    """(f1, f2) of 36.212 table 5.1.3-3 for block size K."""
    f1 = C.c_uint32(0); f2 = C.c_uint32(0)
    lib().op_qpp_params(K, C.byref(f1), C.byref(f2))
    return int(f1.value), int(f2.value)


def qpp_pi(K):
    out = np.zeros(K, np.uint32); lib().op_qpp_pi(K, _p(out)); return out


def turbo_encode(c):
    c = np.ascontiguousarray(c, np.uint8); K = c.size
    out = np.zeros(3 * (K + 4), np.uint8); lib().op_turbo_encode(_p(c), K, _p(out)); return out


def turbo_decode_ref(d, want_dbg=False):
    d = np.ascontiguousarray(d, np.float32).copy(); K = d.size // 3 - 4
    out = np.zeros(K, np.uint8); dbg = np.zeros((6, K), np.int8) if want_dbg else None
    lib().op_turbo_decode_ref(_p(d), K, _p(out), _p(dbg))
    return (out, dbg) if want_dbg else out


def turbo_decode_mlm(d, n_iter=8):
    d = np.ascontiguousarray(d, np.float32); K = d.size // 3 - 4
    out = np.zeros(K, np.uint8); lib().op_turbo_decode_mlm(_p(d), K, n_iter, _p(out)); return out


def turbo_decode_mlm_float(d, n_iter=8):
    d = np.ascontiguousarray(d, np.float32); K = d.size // 3 - 4
    out = np.zeros(K, np.uint8); lib().op_turbo_decode_mlm_float(_p(d), K, n_iter, _p(out)); return out


def siso(a, b):
    a = np.ascontiguousarray(a, np.int8); b = np.ascontiguousarray(b, np.int8)
    out = np.zeros(a.size, np.int8); lib().op_siso(_p(a), _p(b), int(a.size), _p(out)); return out


def rate_unmatch_turbo(e, D, rv=0, C=1, tx_mode=1, N_soft=250368, M_dl_harq=8, chan_type=0):
    e = np.ascontiguousarray(e, np.float32); out = np.full(3 * D, 10000.0, np.float32)
    lib().op_rate_unmatch_turbo(_p(e), int(e.size), D, C, tx_mode, N_soft, M_dl_harq, chan_type, rv, _p(out))
    return out


def rate_match_turbo(d, E, rv=0, C=1, tx_mode=1, N_soft=250368, M_dl_harq=8, chan_type=0):
    d = np.ascontiguousarray(d, np.uint8); D = d.size // 3; out = np.zeros(E, np.uint8)
    lib().op_rate_match_turbo(_p(d), D, C, tx_mode, N_soft, M_dl_harq, chan_type, rv, E, _p(out))
    return out


def prs_c(c_init, n):
    out = np.zeros(n, np.uint8); lib().op_prs_c(C.c_uint32(c_init), n, _p(out)); return out


def crc(bits, poly=0x1864CFB, n_p=24):
    b = np.ascontiguousarray(bits, np.uint8); return int(lib().op_crc(_p(b), int(b.size), C.c_uint32(poly), n_p))


def segmentation_params(tbs):
    K = C.c_int(0); F = C.c_int(0); Cc = C.c_int(0)
    lib().op_segmentation_params(tbs, C.byref(K), C.byref(F), C.byref(Cc)); return K.value, F.value, Cc.value


def zf_1ant(y_re, y_im, h_re, h_im):
    a = [np.ascontiguousarray(v, np.float32) for v in (y_re, y_im, h_re, h_im)]
    xr = np.zeros_like(a[0]); xi = np.zeros_like(a[0])
    lib().op_zf_1ant(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), int(a[0].size), _p(xr), _p(xi)); return xr, xi


def demap(d_re, d_im, mod):
    r = np.ascontiguousarray(d_re, np.float32); i = np.ascontiguousarray(d_im, np.float32)
    out = np.zeros(6 * r.size, np.int8); n = lib().op_demap(_p(r), _p(i), int(r.size), mod, _p(out)); return out[:n]


def noise_var_crs(h_re, h_im, vshift):
    a = np.ascontiguousarray(h_re, np.float32); b = np.ascontiguousarray(h_im, np.float32)
    lib().op_noise_var_crs.restype = C.c_float
    return float(lib().op_noise_var_crs(_p(a), _p(b), int(a.size), int(vshift)))


def soft_bits_1ant(y_re, y_im, h_re, h_im, mod, s2, cseq):
    a = [np.ascontiguousarray(v, np.float32) for v in (y_re, y_im, h_re, h_im)]
    qm = {0: 1, 1: 2, 2: 4, 3: 6}[mod]; M = a[0].size
    cs = np.ascontiguousarray(cseq[:M * qm], np.uint8); out = np.zeros(M * qm, np.int8)
    lib().op_soft_bits_1ant(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), M, mod, C.c_float(s2), _p(cs), _p(out)); return out


def soft_bits_2ant(d_re, d_im, h0_re, h0_im, h1_re, h1_im, mod, s2, cseq):
    a = [np.ascontiguousarray(v, np.float32) for v in (d_re, d_im, h0_re, h0_im, h1_re, h1_im)]
    qm = {0: 1, 1: 2, 2: 4, 3: 6}[mod]; M = a[0].size
    cs = np.ascontiguousarray(cseq[:M * qm], np.uint8); out = np.zeros(M * qm, np.int8)
    lib().op_soft_bits_2ant(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), M, mod, C.c_float(s2), _p(cs), _p(out)); return out


def pdsch_re_indices(N_rb_dl, N_id_cell, N_ant, subfr, cfi, prb, prb1=None):
    p0 = np.ascontiguousarray(prb, np.int32); p1 = p0 if prb1 is None else np.ascontiguousarray(prb1, np.int32)
    out = np.zeros(14 * 12 * p0.size, np.int32)
    n = lib().op_pdsch_re_indices(N_rb_dl, N_id_cell, N_ant, subfr, cfi, int(p0.size), _p(p0), _p(p1), _p(out)); return out[:n]


def fft_forward(re, im):
    re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32)
    o_re = np.zeros_like(re); o_im = np.zeros_like(im)
    lib().op_fft_forward(int(re.size), _p(re), _p(im), _p(o_re), _p(o_im)); return o_re, o_im


def freq_shift(i_buf, q_buf, start_idx, num_samps, freq_offset, fs):
    i_buf = np.ascontiguousarray(i_buf, np.float32).copy(); q_buf = np.ascontiguousarray(q_buf, np.float32).copy()
    lib().op_freq_shift(_p(i_buf), _p(q_buf), C.c_uint32(start_idx), C.c_uint32(num_samps), C.c_float(freq_offset), C.c_uint32(fs))
    return i_buf, q_buf


def crs(N_s, L, N_id_cell):
    re = np.zeros(220, np.float32); im = np.zeros(220, np.float32); lib().op_crs(N_s, L, N_id_cell, _p(re), _p(im)); return re, im


def chan_est(rx_re, rx_im, N_rb_dl, subfr, N_id_cell, N_ant):
    rx_re = np.ascontiguousarray(rx_re, np.float32); rx_im = np.ascontiguousarray(rx_im, np.float32)
    ce_re = np.zeros((4, 16, 1200), np.float32); ce_im = np.zeros((4, 16, 1200), np.float32)
    lib().op_chan_est(_p(rx_re), _p(rx_im), N_rb_dl, subfr, N_id_cell, N_ant, _p(ce_re), _p(ce_im)); return ce_re, ce_im


def get_dl_subframe_and_ce(i_samps, q_samps, fs_idx, N_rb_dl, frame_start_idx, subfr, N_id_cell, N_ant):
    rx_re = np.zeros((16, 1200), np.float32); rx_im = np.zeros((16, 1200), np.float32)
    ce_re = np.zeros((4, 16, 1200), np.float32); ce_im = np.zeros((4, 16, 1200), np.float32)
    lib().op_get_dl_subframe_and_ce(_p(i_samps), _p(q_samps), fs_idx, N_rb_dl, C.c_uint32(frame_start_idx), subfr, N_id_cell, N_ant,
                                    _p(rx_re), _p(rx_im), _p(ce_re), _p(ce_im))
    return rx_re, rx_im, ce_re, ce_im


def pdsch_decode(rx_re, rx_im, ce_re, ce_im, N_rb_dl, N_id_cell, subfr, cfi, mod, prb, tbs, rnti, rv=0, turbo_mode=0, n_iter=8):
    prb = np.ascontiguousarray(prb, np.int32); out = np.zeros(tbs + 32, np.uint8)
    a = [np.ascontiguousarray(v, np.float32) for v in (rx_re, rx_im, ce_re, ce_im)]
    rc = lib().op_pdsch_decode(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), N_rb_dl, N_id_cell, subfr, cfi, mod, int(prb.size), _p(prb), tbs, rv,
                               rnti, turbo_mode, n_iter, _p(out))
    return rc, out[:tbs]


def find_coarse_timing(i_s, q_s, fs_idx, N_slots):
    fo = np.zeros(5, np.float32); ss = np.zeros((5, 7), np.uint32); n = C.c_int(0)
    lib().op_find_coarse_timing(_p(i_s), _p(q_s), fs_idx, N_slots, _p(fo), _p(ss), C.byref(n))
    return fo, ss, n.value


def find_pss(i_s, q_s, fs_idx, N_rb_dl, symb_starts):
    ss = np.ascontiguousarray(symb_starts, np.uint32).copy()
    a = C.c_uint32(0); b = C.c_uint32(0); th = C.c_float(0); fo = C.c_float(0)
    lib().op_find_pss(_p(i_s), _p(q_s), fs_idx, N_rb_dl, _p(ss), C.byref(a), C.byref(b), C.byref(th), C.byref(fo))
    return ss, a.value, b.value, th.value, fo.value


def find_sss(i_s, q_s, fs_idx, N_rb_dl, N_id_2, symb_starts, pss_thresh):
    ss = np.ascontiguousarray(symb_starts, np.uint32).copy()
    a = C.c_uint32(0); b = C.c_uint32(0)
    rc = lib().op_find_sss(_p(i_s), _p(q_s), fs_idx, N_rb_dl, C.c_uint32(N_id_2), _p(ss), C.c_float(pss_thresh), C.byref(a), C.byref(b))
    return rc, ss, a.value, b.value


def pre_decoder_dl(y_re, y_im, h_re, h_im, N_ant):
    a = [np.ascontiguousarray(v, np.float32) for v in (y_re, y_im, h_re, h_im)]
    M = a[0].size; d_re = np.zeros(M + 4, np.float32); d_im = np.zeros(M + 4, np.float32)
    n = lib().op_pre_decoder_dl(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), M, M, N_ant, _p(d_re), _p(d_im))
    return d_re[:n], d_im[:n]


def pdsch_decode_nant(rx_re, rx_im, ce_re, ce_im, N_rb_dl, N_id_cell, N_ant, subfr, cfi, mod, prb, tbs, rnti, rv=0, turbo_mode=0, n_iter=8):
    """ce_*: [N_ant][16][1200]; returns (rc, bits, descrambled soft bits)"""
    prb = np.ascontiguousarray(prb, np.int32); out = np.zeros(tbs + 32, np.uint8)
    a = [np.ascontiguousarray(v, np.float32) for v in (rx_re, rx_im, ce_re, ce_im)]
    soft = np.zeros(14 * 12 * prb.size * 6, np.float32); n = C.c_int(0)
    rc = lib().op_pdsch_decode_nant(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), N_rb_dl, N_id_cell, N_ant, subfr, cfi, mod, int(prb.size), _p(prb), tbs, rv,
                                    rnti, turbo_mode, n_iter, _p(out), _p(soft), C.byref(n))
    return rc, out[:tbs], soft[:n.value]


# ---- control channels (oracle/port/control.cc)
def _alloc_dict(o):
    n = int(o[3])
    return dict(rnti=int(o[0]), tbs=int(o[1]), mcs=int(o[2]), N_prb=n, rv_idx=int(o[4]), mod_type=int(o[5]), tx_mode=int(o[6]),
                chan_type=int(o[7]), prb=[int(x) for x in o[8:8 + min(n, 110)]], prb1=[int(x) for x in o[118:118 + min(n, 110)]])


def tbs_1a(i_tbs, n_prb_1a):
    return lib().op_tbs_1a(i_tbs, n_prb_1a)


def tbs_1c(i):
    return lib().op_tbs_1c(i)


def rate_unmatch_conv(e, N_c):
    e = np.ascontiguousarray(e, np.float32); d = np.zeros(3 * N_c, np.float32)
    lib().op_rate_unmatch_conv(_p(e), int(e.size), N_c, _p(d))
    return d


def viterbi_decode(d):
    d = np.ascontiguousarray(d, np.float32); c = np.zeros(d.size + 80, np.uint8)
    n = lib().op_viterbi_decode(_p(d), int(d.size), _p(c))
    return c[:n].copy()


def bch_channel_decode(rx_re, rx_im, ce_re, ce_im, N_rb_dl, N_id_cell):
    """ce_*: [4][16][1200].  Returns (err, N_ant, bits24, offset)."""
    rx_re, rx_im, ce_re, ce_im = [np.ascontiguousarray(a, np.float32) for a in (rx_re, rx_im, ce_re, ce_im)]
    na = C.c_int(0); off = C.c_int(0); out = np.zeros(24, np.uint8)
    err = lib().op_bch_channel_decode(_p(rx_re), _p(rx_im), _p(ce_re), _p(ce_im), N_rb_dl, N_id_cell, C.byref(na), _p(out), C.byref(off))
    return err, na.value, out, off.value


def pdcch_cce_map(N_rb_dl, N_id_cell, N_ant, N_symbs, phich_res=1.0):
    re_of = np.zeros((N_symbs * N_rb_dl * 3 // 9 + 1) * 36, np.int32); nreg = C.c_int(0)
    n = lib().op_pdcch_cce_map(N_rb_dl, N_id_cell, N_ant, N_symbs, C.c_float(phich_res), _p(re_of), C.byref(nreg))
    return re_of[:n * 36].reshape(n, 36).copy(), nreg.value


def dci_1a_unpack(bits, rnti, N_rb_dl, N_ant):
    b = np.ascontiguousarray(bits, np.uint8); o = np.zeros(228, np.int32)
    err = lib().op_dci_1a_unpack(_p(b), rnti, N_rb_dl, N_ant, _p(o))
    return err, _alloc_dict(o)


def dci_1c_unpack(bits, rnti, N_rb_dl, N_ant):
    b = np.ascontiguousarray(bits, np.uint8); o = np.zeros(228, np.int32)
    err = lib().op_dci_1c_unpack(_p(b), rnti, N_rb_dl, N_ant, _p(o))
    return err, _alloc_dict(o)


def pdcch_channel_decode(rx_re, rx_im, ce_re, ce_im, N_rb_dl, N_id_cell, N_ant, subfr, phich_res=1.0):
    """Returns dict(err, cfi, N_symbs, allocs, raw_found[12], raw_rnti[12], raw_bits[12][32])."""
    rx_re, rx_im, ce_re, ce_im = [np.ascontiguousarray(a, np.float32) for a in (rx_re, rx_im, ce_re, ce_im)]
    cfi = C.c_int(0); ns = C.c_int(0); na = C.c_int(0); al = np.zeros((6, 228), np.int32)
    rf = np.zeros(12, np.int32); rr = np.zeros(12, np.int32); rb = np.zeros((12, 32), np.uint8)
    err = lib().op_pdcch_channel_decode(_p(rx_re), _p(rx_im), _p(ce_re), _p(ce_im), N_rb_dl, N_id_cell, N_ant, subfr, C.c_float(phich_res),
                                        C.byref(cfi), C.byref(ns), C.byref(na), _p(al), _p(rf), _p(rr), _p(rb))
    return dict(err=err, cfi=cfi.value, N_symbs=ns.value, allocs=[_alloc_dict(al[a]) for a in range(na.value)], raw_found=rf, raw_rnti=rr,
                raw_bits=rb)

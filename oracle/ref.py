"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libref_oracle.so.

That library is the reference's own liblte_phy.cc compiled unmodified (see
oracle/Makefile, oracle/ref_wrap.cc).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")

FS = {1.92: 0, 3.84: 1, 7.68: 2, 15.36: 3, 30.72: 4}
MOD_BPSK, MOD_QPSK, MOD_16QAM, MOD_64QAM = 0, 1, 2, 3
RX_NULL = 10000.0


def available():
    return os.path.exists(_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_oracle.so missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(_SO)
        _lib.ref_phy_init.restype = C.c_void_p
        _lib.ref_phy_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        _lib.ref_subframe_new.restype = C.c_void_p
        _lib.ref_subframe_field.restype = C.POINTER(C.c_float)
        _lib.ref_subframe_field.argtypes = [C.c_void_p, C.c_int]
        _lib.ref_pdsch_scratch.restype = C.c_void_p
        _lib.ref_pdsch_scratch.argtypes = [C.c_void_p, C.c_int]
        _lib.ref_timing_abs_corr.restype = C.POINTER(C.c_float)
        _lib.ref_timing_abs_corr.argtypes = [C.c_void_p]
        _lib.ref_sizeof_phy_struct.restype = C.c_ulong
        _lib.ref_sizeof_subframe_struct.restype = C.c_ulong
    return _lib


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class Subframe:
    """Owns one LIBLTE_PHY_SUBFRAME_STRUCT (liblte_phy.h:226-239) and exposes numpy views."""

    def __init__(self):
        self.h = C.c_void_p(lib().ref_subframe_new())

    def __del__(self):
        try:
            lib().ref_subframe_free(self.h)
        except Exception:
            pass

    def clear(self):
        lib().ref_subframe_clear(self.h)

    def set_num(self, n):
        lib().ref_subframe_set_num(self.h, int(n))

    def _view(self, field, shape):
        ptr = lib().ref_subframe_field(self.h, field)
        return np.ctypeslib.as_array(ptr, shape=shape)

    @property
    def rx_symb_re(self): return self._view(0, (16, 1200))
    @property
    def rx_symb_im(self): return self._view(1, (16, 1200))
    @property
    def rx_ce_re(self): return self._view(2, (4, 16, 1200))
    @property
    def rx_ce_im(self): return self._view(3, (4, 16, 1200))
    @property
    def tx_symb_re(self): return self._view(4, (4, 16, 1200))
    @property
    def tx_symb_im(self): return self._view(5, (4, 16, 1200))


class RefPhy:
    """One LIBLTE_PHY_STRUCT of the reference (zero-initialised)."""

    def __init__(self, fs=30.72, N_id_cell=0, N_ant=1, N_rb_dl=100, phich_res=1.0):
        self.h = C.c_void_p(lib().ref_phy_init(FS[fs], N_id_cell, N_ant, N_rb_dl, phich_res))
        if not self.h:
            raise RuntimeError("liblte_phy_init failed")
        self.N_rb_dl = N_rb_dl

    def close(self):
        if self.h:
            lib().ref_phy_cleanup(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_n_rb_dl(self, n):
        self.N_rb_dl = n
        return lib().ref_phy_update_n_rb_dl(self.h, int(n))

    # ---- turbo
    def turbo_encode(self, c_bits):
        c = _u8(c_bits); K = c.size
        out = np.zeros(3 * (K + 4), np.uint8)
        n = lib().ref_turbo_encode(self.h, _p(c), K, _p(out))
        assert n == 3 * (K + 4)
        return out  # branch-major d[x*(K+4)+i]

    def turbo_decode(self, d_bits):
        d = _f32(d_bits).copy(); K = d.size // 3 - 4
        out = np.zeros(K + 16, np.uint8)
        lib().ref_turbo_decode(self.h, _p(d), int(d.size), _p(out))
        return out[:K]

    def turbo_intermediate(self, which, K):
        out = np.zeros(K, np.int8)
        lib().ref_turbo_get_intermediate(self.h, which, K, _p(out))
        return out

    def viterbi_siso(self, d_bits):
        d = np.ascontiguousarray(d_bits, np.int8)
        out = np.zeros(d.size // 2 + 16, np.int8); n = C.c_int(0)
        lib().ref_viterbi_decode_siso(self.h, _p(d), int(d.size), _p(out), C.byref(n))
        return out[:d.size // 2]

    # ---- rate matching
    def rate_match_turbo(self, d_bits, E, rv=0, N_codeblocks=1, tx_mode=1, N_soft=250368, M_dl_harq=8, chan_type=0):
        d = _u8(d_bits); out = np.zeros(E, np.uint8)
        lib().ref_rate_match_turbo(self.h, _p(d), int(d.size), N_codeblocks, tx_mode, N_soft, M_dl_harq, chan_type, rv, E, _p(out))
        return out

    def rate_unmatch_turbo(self, e_bits, D, rv=0, N_codeblocks=1, tx_mode=1, N_soft=250368, M_dl_harq=8, chan_type=0):
        e = _f32(e_bits); out = np.zeros(3 * D + 96, np.float32)
        n = lib().ref_rate_unmatch_turbo(self.h, _p(e), int(e.size), D, N_codeblocks, tx_mode, N_soft, M_dl_harq, chan_type, rv, _p(out))
        return out[:n]

    # ---- PDSCH chain
    def dlsch_encode(self, tb_bits, G, Q_m, rv=0, tx_mode=1):
        tb = _u8(tb_bits); out = np.zeros(G + 64, np.uint8)
        n = lib().ref_dlsch_channel_encode(self.h, _p(tb), int(tb.size), tx_mode, rv, G, Q_m, _p(out))
        return out[:n]

    def dlsch_decode(self, llr, tbs, rv=0, tx_mode=1):
        x = _f32(llr); out = np.zeros(tbs + 64, np.uint8); n = C.c_int(0)
        err = lib().ref_dlsch_channel_decode(self.h, _p(x), int(x.size), tbs, tx_mode, rv, _p(out), C.byref(n))
        return err, out[:tbs]

    def get_dl_subframe_and_ce(self, i_samps, q_samps, frame_start_idx, subfr_num, N_id_cell, N_ant, sf):
        return lib().ref_get_dl_subframe_and_ce(self.h, _p(i_samps), _p(q_samps), C.c_uint(frame_start_idx), subfr_num,
                                                N_id_cell, N_ant, sf.h)

    def samples_to_symbols_dl(self, i_samps, q_samps, slot_start_idx, symbol_offset):
        n = 12 * self.N_rb_dl
        re = np.zeros(1200, np.float32); im = np.zeros(1200, np.float32)
        lib().ref_samples_to_symbols_dl(self.h, _p(i_samps), _p(q_samps), C.c_uint(slot_start_idx), symbol_offset, _p(re), _p(im))
        return re[:n], im[:n]

    def pdsch_encode(self, sf, cfi, N_id_cell, N_ant, mod, prb, tbs, rnti, tb_bits, rv=0, tx_mode=1):
        prb = np.ascontiguousarray(prb, np.int32); tb = _u8(tb_bits)
        return lib().ref_pdsch_channel_encode(self.h, sf.h, cfi, N_id_cell, N_ant, mod, int(prb.size), _p(prb), tbs, rv, rnti,
                                              tx_mode, _p(tb))

    def pdsch_decode(self, sf, cfi, N_id_cell, N_ant, mod, prb, tbs, rnti, rv=0, tx_mode=1):
        prb = np.ascontiguousarray(prb, np.int32)
        out = np.zeros(tbs + 64, np.uint8); n = C.c_int(0)
        err = lib().ref_pdsch_channel_decode(self.h, sf.h, cfi, N_id_cell, N_ant, mod, int(prb.size), _p(prb), tbs, rv, rnti,
                                             tx_mode, _p(out), C.byref(n))
        return err, out[:tbs]

    def pdsch_decode_big(self, sf, cfi, N_id_cell, mod, prb, tbs, rnti, rv=0, tx_mode=1, want_soft=False):
        prb = np.ascontiguousarray(prb, np.int32)
        out = np.zeros(tbs + 64, np.uint8); n = C.c_int(0)
        soft = np.zeros(12 * 14 * prb.size * 6, np.int8) if want_soft else None
        err = lib().ref_pdsch_decode_big(self.h, sf.h, cfi, N_id_cell, mod, int(prb.size), _p(prb), tbs, rv, rnti, tx_mode,
                                         _p(out), C.byref(n), _p(soft) if want_soft else None)
        return (err, out[:tbs], soft) if want_soft else (err, out[:tbs])

    def pdsch_scratch(self, which, n, dtype):
        ptr = lib().ref_pdsch_scratch(self.h, which)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()

    # ---- TX helpers
    def map_crs(self, sf, N_id_cell, N_ant): return lib().ref_map_crs(self.h, sf.h, N_id_cell, N_ant)
    def map_pss(self, sf, N_id_2, N_ant): return lib().ref_map_pss(self.h, sf.h, N_id_2, N_ant)
    def map_sss(self, sf, N_id_1, N_id_2, N_ant): return lib().ref_map_sss(self.h, sf.h, N_id_1, N_id_2, N_ant)

    def create_dl_subframe(self, sf, ant, n_samps):
        i = np.zeros(n_samps, np.float32); q = np.zeros(n_samps, np.float32)
        lib().ref_create_dl_subframe(self.h, sf.h, ant, _p(i), _p(q))
        return i, q

    # ---- sync
    def find_coarse_timing(self, i_samps, q_samps, N_slots):
        fo = np.zeros(5, np.float32); ss = np.zeros((5, 7), np.uint32); n = C.c_int(0)
        err = lib().ref_find_coarse_timing(self.h, _p(i_samps), _p(q_samps), N_slots, _p(fo), _p(ss), C.byref(n))
        return err, fo, ss, n.value

    def timing_abs_corr(self, n):
        return np.ctypeslib.as_array(lib().ref_timing_abs_corr(self.h), shape=(n,)).copy()

    def find_pss(self, i_samps, q_samps, symb_starts):
        ss = np.ascontiguousarray(symb_starts, np.uint32).copy()
        nid2 = C.c_uint(0); ps = C.c_uint(0); th = C.c_float(0); fo = C.c_float(0)
        err = lib().ref_find_pss(self.h, _p(i_samps), _p(q_samps), _p(ss), C.byref(nid2), C.byref(ps), C.byref(th), C.byref(fo))
        return err, ss, nid2.value, ps.value, th.value, fo.value

    def find_sss(self, i_samps, q_samps, N_id_2, symb_starts, pss_thresh):
        ss = np.ascontiguousarray(symb_starts, np.uint32).copy()
        nid1 = C.c_uint(0); fsi = C.c_uint(0)
        err = lib().ref_find_sss(self.h, _p(i_samps), _p(q_samps), C.c_uint(N_id_2), _p(ss), C.c_float(pss_thresh), C.byref(nid1), C.byref(fsi))
        return err, ss, nid1.value, fsi.value

    # ---- control channels
    def bch_decode(self, sf, N_id_cell):
        na = C.c_int(0); n = C.c_int(0); off = C.c_int(0); out = np.zeros(64, np.uint8)
        err = lib().ref_bch_channel_decode(self.h, sf.h, N_id_cell, C.byref(na), _p(out), C.byref(n), C.byref(off))
        return err, na.value, out[:n.value], off.value

    def bch_encode(self, sf, bits, N_id_cell, N_ant, sfn):
        b = _u8(bits)
        return lib().ref_bch_channel_encode(self.h, sf.h, _p(b), int(b.size), N_id_cell, N_ant, C.c_uint(sfn))

    def pdcch_encode(self, sf, N_id_cell, N_ant, cfi, allocs, phich_res=1.0, phich_dur=0):
        """allocs: list of dict(rnti, mcs, rv_idx, N_prb, prb0) -- DCI 1A in the common search space, aggregation level 4"""
        a = np.array([[x["rnti"], x["mcs"], x.get("rv_idx", 0), x["N_prb"], x["prb0"]] for x in allocs], np.int32).reshape(-1, 5)
        return lib().ref_pdcch_channel_encode(self.h, sf.h, N_id_cell, N_ant, C.c_float(phich_res), phich_dur, cfi, len(allocs), _p(a))

    def pdcch_decode2(self, sf, N_id_cell, N_ant, phich_res=1.0, phich_dur=0):
        cfi = C.c_int(0); ns = C.c_int(0); na = C.c_int(0); out = np.zeros((6, 228), np.int32)
        err = lib().ref_pdcch_channel_decode2(self.h, sf.h, N_id_cell, N_ant, C.c_float(phich_res), phich_dur, C.byref(cfi), C.byref(ns),
                                              C.byref(na), _p(out))
        return err, cfi.value, ns.value, [_alloc_dict(out[a]) for a in range(na.value)]

    def rate_unmatch_conv(self, e_bits, N_c):
        e = _f32(e_bits); d = np.zeros(3 * N_c + 8, np.float32); n = C.c_int(0)
        lib().ref_rate_unmatch_conv(self.h, _p(e), int(e.size), N_c, _p(d), C.byref(n))
        return d[:n.value].copy()

    def viterbi_decode(self, d_bits):
        d = _f32(d_bits); c = np.zeros(d.size + 80, np.uint8); n = C.c_int(0)
        lib().ref_viterbi_decode(self.h, _p(d), int(d.size), _p(c), C.byref(n))
        return c[:n.value].copy()

    def pdcch_decode(self, sf, N_id_cell, N_ant, phich_res=1.0, phich_dur=0):
        cfi = C.c_int(0); na = C.c_int(0); out = np.zeros((6, 118), np.int32)
        err = lib().ref_pdcch_channel_decode(self.h, sf.h, N_id_cell, N_ant, C.c_float(phich_res), phich_dur, C.byref(cfi), C.byref(na), _p(out))
        allocs = []
        for a in range(na.value):
            o = out[a]
            allocs.append(dict(rnti=int(o[0]), tbs=int(o[1]), mcs=int(o[2]), N_prb=int(o[3]), rv_idx=int(o[4]), mod_type=int(o[5]),
                               tx_mode=int(o[6]), chan_type=int(o[7]), prb=[int(x) for x in o[8:8 + o[3]]]))
        return err, cfi.value, allocs


_native = None


def native_lib():
    """The -march=native build of the same reference sources (oracle/Makefile), loaded only where the host CPU has every flag of the
    build host; otherwise the portable build.  Used by bench.py's CPU arm only."""
    global _native
    if _native is None:
        so = os.path.join(_HERE, "_ref", "libref_oracle_native.so")
        fl = os.path.join(_HERE, "_ref", "native_flags.txt")
        ok = False
        if os.path.exists(so) and os.path.exists(fl):
            try:
                need = set(open(fl).read().split())
                have = set()
                for ln in open("/proc/cpuinfo"):
                    if ln.startswith("flags"):
                        have = set(ln.split(":", 1)[1].split())
                        break
                ok = need <= have
            except OSError:
                ok = False
        v4 = os.path.join(_HERE, "_ref", "libref_oracle_v4.so")
        have_v4 = False
        try:
            for ln in open("/proc/cpuinfo"):
                if ln.startswith("flags"):
                    have_v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= set(ln.split(":", 1)[1].split())
                    break
        except OSError:
            pass
        if ok:
            _native = (C.CDLL(so), "-O3 -march=native")
        elif have_v4 and os.path.exists(v4):
            _native = (C.CDLL(v4), "-O3 -march=x86-64-v4")
        else:
            _native = (lib(), "-O3 -march=x86-64-v3")
        _native[0].ref_chain_pool.restype = C.c_int
    return _native


def chain_pool(n_threads, i_samps, q_samps, start, subfr, allocs, cell, cfi, n_rb=100, fs=30.72, want_tb=None, seg=1):
    """Native std::thread pool over subframes (ref_wrap.cc: ref_chain_pool): the reference's liblte_phy_get_dl_subframe_and_ce + PDSCH
    receive of every allocation.  Returns dict(seconds, fails, front_s, pdsch_s, flags)."""
    L, flags = native_lib()
    n_sf = len(start)
    desc = np.zeros((len(allocs), 114), np.int32)
    for a, al in enumerate(allocs):
        desc[a, 0] = al["mod"]; desc[a, 1] = al["tbs"]; desc[a, 2] = al["rnti"]; desc[a, 3] = len(al["prb"]); desc[a, 4:4 + len(al["prb"])] = al["prb"]
    st = np.ascontiguousarray(start, np.int64); sn = np.ascontiguousarray(subfr, np.int32)
    sec = C.c_double(0); stage = (C.c_double * 2)()
    if want_tb is not None:
        want_tb = np.ascontiguousarray(want_tb, np.uint8)
        wt, stride = _p(want_tb), int(want_tb.shape[-1])
    else:
        wt, stride = None, 0
    fails = L.ref_chain_pool(int(n_threads), FS[fs], int(n_rb), int(cell), int(cfi), _p(i_samps), _p(q_samps), n_sf, _p(st), _p(sn),
                             len(allocs), _p(desc), wt, int(seg), stride, C.byref(sec), stage)
    if fails < 0:
        raise RuntimeError("liblte_phy_init failed inside ref_chain_pool")
    return dict(seconds=sec.value, fails=int(fails), front_s=stage[0], pdsch_s=stage[1], flags=flags)


def prs_c(c_init, n):
    out = np.zeros(n, np.uint32)
    lib().ref_generate_prs_c(C.c_uint(c_init), n, _p(out))
    return out


def crs(N_s, L, N_id_cell):
    re = np.zeros(220, np.float32); im = np.zeros(220, np.float32)
    lib().ref_generate_crs(N_s, L, N_id_cell, _p(re), _p(im))
    return re, im


def crc(bits, which=0):
    a = _u8(bits); out = np.zeros(24 if which < 2 else 16, np.uint8)
    lib().ref_calc_crc(_p(a), int(a.size), which, _p(out))
    return out


def mod_map(bits, mod):
    b = _u8(bits); re = np.zeros(b.size, np.float32); im = np.zeros(b.size, np.float32)
    m = lib().ref_modulation_mapper(_p(b), int(b.size), mod, _p(re), _p(im))
    return re[:m], im[:m]


def mod_demap(d_re, d_im, mod):
    r = _f32(d_re); i = _f32(d_im); out = np.zeros(r.size * 6, np.int8)
    n = lib().ref_modulation_demapper(_p(r), _p(i), int(r.size), mod, _p(out))
    return out[:n]


def pre_decoder(y_re, y_im, h_re, h_im, N_ant=1):
    y_re = _f32(y_re); y_im = _f32(y_im); h_re = _f32(h_re); h_im = _f32(h_im)
    M = y_re.size; h_len = h_re.size // N_ant
    x_re = np.zeros(M + 8, np.float32); x_im = np.zeros(M + 8, np.float32)
    ml = lib().ref_pre_decoder_dl(_p(y_re), _p(y_im), _p(h_re), _p(h_im), h_len, M, N_ant, _p(x_re), _p(x_im))
    return x_re[:M], x_im[:M], ml


def _alloc_dict(o):
    n = int(o[3])
    return dict(rnti=int(o[0]), tbs=int(o[1]), mcs=int(o[2]), N_prb=n, rv_idx=int(o[4]), mod_type=int(o[5]), tx_mode=int(o[6]),
                chan_type=int(o[7]), prb=[int(x) for x in o[8:8 + n]], prb1=[int(x) for x in o[118:118 + n]])


def dci_1a_unpack(bits, rnti, N_rb_dl, N_ant):
    b = _u8(bits); o = np.zeros(228, np.int32)
    err = lib().ref_dci_1a_unpack(_p(b), int(b.size), rnti, N_rb_dl, N_ant, _p(o))
    return err, _alloc_dict(o)


def dci_1c_unpack(bits, rnti, N_rb_dl, N_ant):
    b = _u8(bits); o = np.zeros(228, np.int32)
    err = lib().ref_dci_1c_unpack(_p(b), int(b.size), rnti, N_rb_dl, N_ant, _p(o))
    return err, _alloc_dict(o)


def tbs_71721(i_tbs, n_prb):
    return lib().ref_tbs_71721(i_tbs, n_prb)


def tbs_71723(i):
    return lib().ref_tbs_71723(i)

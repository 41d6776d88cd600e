"""ctypes loader for libopenlte_b200.so (the C ABI declared in include/lte_b200.h).

Fails loudly when the shared library has not been built: there is no Python or CPU fallback path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libopenlte_b200.so")

OK, ERR_ARGS, ERR_CUDA, ERR_DECODE, ERR_UNSUPPORTED, ERR_CONTENTS = 0, -1, -2, -3, -4, -5
FS_IDX = {1.92: 0, 3.84: 1, 7.68: 2, 15.36: 3, 30.72: 4}
MOD_BPSK, MOD_QPSK, MOD_16QAM, MOD_64QAM = 0, 1, 2, 3
TURBO_REF_EXACT, TURBO_MAXLOGMAP = 0, 1
IQ_F32, IQ_I8 = 0, 1


class LibraryMissing(RuntimeError):
    pass


class Alloc(C.Structure):
    """lteb200_alloc (include/lte_b200.h)."""
    _fields_ = [("mod_type", C.c_int32), ("N_prb", C.c_int32), ("prb", (C.c_int32 * 110) * 2), ("tbs", C.c_int32),
                ("rv_idx", C.c_int32), ("tx_mode", C.c_int32), ("rnti", C.c_int32)]

    @classmethod
    def make(cls, mod_type, prb, tbs, rnti, rv_idx=0, tx_mode=1, prb_slot1=None):
        a = cls()
        a.mod_type = mod_type; a.N_prb = len(prb); a.tbs = tbs; a.rv_idx = rv_idx; a.tx_mode = tx_mode; a.rnti = rnti
        p1 = prb if prb_slot1 is None else prb_slot1
        for j, (x, y) in enumerate(zip(prb, p1)):
            a.prb[0][j] = int(x); a.prb[1][j] = int(y)
        return a


class BchResult(C.Structure):
    """lteb200_bch_result"""
    _fields_ = [("ok", C.c_int32), ("N_ant", C.c_int32), ("offset", C.c_int32), ("bits", C.c_uint8 * 24)]


class PdcchRaw(C.Structure):
    """lteb200_pdcch_raw"""
    _fields_ = [("cfi", C.c_int32), ("N_symbs", C.c_int32), ("found", C.c_int32 * 12), ("rnti", C.c_int32 * 12), ("bits", (C.c_uint8 * 32) * 12)]


class Dci(C.Structure):
    """lteb200_dci"""
    _fields_ = [("alloc", Alloc), ("mcs", C.c_int32), ("format", C.c_int32)]

    def as_dict(self):
        a = self.alloc; n = min(int(a.N_prb), 110)
        return dict(rnti=int(a.rnti), tbs=int(a.tbs), mcs=int(self.mcs), N_prb=int(a.N_prb), rv_idx=int(a.rv_idx), mod_type=int(a.mod_type),
                    tx_mode=int(a.tx_mode), chan_type=0, prb=[int(a.prb[0][j]) for j in range(n)], prb1=[int(a.prb[1][j]) for j in range(n)])


class ScanResult(C.Structure):
    """lteb200_scan_result"""
    _fields_ = [("cfi", C.c_int32), ("N_symbs", C.c_int32), ("N_alloc", C.c_int32), ("pdcch_err", C.c_int32), ("dci", Dci * 6), ("crc_ok", C.c_int32 * 6),
                ("tb_offset", C.c_longlong * 6)]


# name -> (restype, argtypes); every symbol declared in include/lte_b200.h
_vp, _i, _ll, _u32, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_uint32, C.c_float
SYMBOLS = {
    "lteb200_create": (_i, [C.POINTER(_vp), _i]),
    "lteb200_destroy": (None, [_vp]),
    "lteb200_last_error": (C.c_char_p, [_vp]),
    "lteb200_launch_count": (C.c_ulonglong, [_vp]),
    "lteb200_version": (C.c_char_p, []),
    "lteb200_turbo_decode_batch": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "lteb200_turbo_decode": (_i, [_vp, _vp, _u32, _u32, _vp, _i, _i]),
    "lteb200_rate_unmatch_turbo_batch": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "lteb200_rate_unmatch_turbo": (_i, [_vp, _vp, _u32, _vp, _u32, _u32, _u32, _u32, _u32, _i, _u32, _vp, _vp]),
    "lteb200_get_dl_subframe_and_ce_batch": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "lteb200_get_dl_subframe_and_ce": (_i, [_vp, _vp, _vp, _ll, _i, _i, _u32, C.c_uint8, _u32, C.c_uint8, _vp, _vp, _vp, _vp]),
    "lteb200_pdsch_decode_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "lteb200_pdsch_channel_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _u32, _vp, _u32, _u32, C.c_uint8, _i, _i, _vp, _vp]),
    "lteb200_pdsch_demod_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "lteb200_ofdm_demod_batch": (_i, [_vp, _vp, _i, _ll, _i, _i, _vp, _i, _vp, _vp]),
    "lteb200_chan_est_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "lteb200_dl_receive_batch": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "lteb200_dl_receive_host": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp]),
    "lteb200_find_coarse_timing_and_freq_offset": (_i, [_vp, _vp, _vp, _ll, _i, _u32, _vp, _vp, _vp]),
    "lteb200_find_pss_and_fine_timing": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "lteb200_find_sss": (_i, [_vp, _vp, _vp, _ll, _i, _i, _u32, _vp, _f, _vp, _vp]),
    "lteb200_bch_decode_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "lteb200_bch_channel_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _u32, _vp, _vp, _vp, _vp]),
    "lteb200_pdcch_decode_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _i, _vp, _vp]),
    "lteb200_pdcch_cce_map": (_i, [_i, _i, _i, _i, _f, _vp]),
    "lteb200_pdsch_re_indices": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "lteb200_rate_unmatch_map": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "lteb200_e2e_plan": (_i, [_i, _i, _i, _i, _i, _i, _ll, _vp, _vp, _vp, _i, _vp]),
    "lteb200_pdcch_unpack": (_i, [_vp, _i, _i, _vp, _vp]),
    "lteb200_pdcch_channel_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _u32, _u32, C.c_uint8, _f, _i, _vp, _vp, _vp, _vp]),
    "lteb200_dl_scan_batch": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _f, _vp, _vp, _i, _i, _i, _vp, _vp, _ll, _vp, _vp]),
    "lteb200_set_demapper": (_i, [_vp, _i]),
    "lteb200_set_tb_format": (_i, [_vp, _i]),
    "lteb200_set_interleaver": (_i, [_vp, _i]),
    "lteb200_q8_stride": (C.c_size_t, [_i]),
    "lteb200_pdsch_llr_q8_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "lteb200_turbo_decode_q8_batch": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "lteb200_dci_1a_tbs": (_i, [_i]),
    "lteb200_dci_1a_encode": (_i, [_vp, _i, _i, _vp]),
    "lteb200_bch_encode": (_i, [_vp, _i, _i, _vp]),
    "lteb200_dl_transmit_frames_batch": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "lteb200_dl_transmit_batch": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "lteb200_freq_shift_batch": (_i, [_vp, _vp, _i, _ll, _ll, _f, _i, _vp, _vp]),
    "lteb200_freq_shift": (_i, [_vp, _vp, _vp, _u32, _u32, _f, _i]),
    "lteb200_set_freq_offset": (_i, [_vp, _f, _ll]),
    "lteb200_prs_c_batch": (_i, [_vp, _u32, _i, _vp, _vp]),
    "lteb200_crc24a_check_batch": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
}


class TxControl(C.Structure):
    """lteb200_tx_control (include/lte_b200.h)"""
    _fields_ = [("N_id_cell", C.c_int32), ("N_ant", C.c_int32), ("ant_gain_re", C.c_float * 4), ("ant_gain_im", C.c_float * 4), ("sync", C.c_int32),
                ("pbch", C.c_int32), ("sfn_first", C.c_int32), ("mib_bits", C.c_void_p), ("pdcch", C.c_int32), ("phich_res", C.c_float), ("mcs", C.c_int32 * 8)]


class Lib:
    def __init__(self, path=SO_PATH):
        if not os.path.exists(path):
            raise LibraryMissing(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C openlte_b200/csrc`). openlte_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.dll, name)   # AttributeError if the header promises a symbol the library lacks
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[len("lteb200_"):], fn)

    def new_ctx(self, device=-1):
        h = _vp()
        rc = self.create(C.byref(h), device)
        if rc != OK:
            raise RuntimeError(f"lteb200_create failed ({rc}): no usable CUDA device; there is no CPU fallback")
        return Ctx(self, h)


class Ctx:
    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def close(self):
        if self.h:
            self.lib.destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, allow=()):
        if rc != OK and rc not in allow:
            raise RuntimeError(f"lteb200 error {rc}: {self.lib.last_error(self.h).decode()}")
        return rc

    @property
    def launches(self):
        return int(self.lib.launch_count(self.h))


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = Lib()
    return _cached

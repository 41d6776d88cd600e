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

"""Cell search (coarse timing, PSS, SSS) on the reference's capture: wall time of the host-pointer twins on one GPU beside the reference's
own functions on one host core.  The scan runs these once per acquisition, on 160 slots (LTE_fdd_dl_fs_samp_buf.cc:285).  Reporting aid."""
import sys, os, json, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import openlte_b200

z = np.load(os.path.join(ROOT, "tests", "golden", "capture_cfg1.npz"))
i_s = np.ascontiguousarray(z["iq"][:, 0].astype(np.float32)); q_s = np.ascontiguousarray(z["iq"][:, 1].astype(np.float32))
L = openlte_b200.load(); ctx = L.new_ctx(0)
p = lambda a: a.ctypes.data_as(C.c_void_p)
N_slots = 20


def gpu_once():
    fo = np.zeros(5, np.float32); ss = np.zeros((5, 7), np.uint32); n = C.c_uint32(0)
    t0 = time.perf_counter()
    ctx.check(L.find_coarse_timing_and_freq_offset(ctx.h, p(i_s), p(q_s), i_s.size, 4, N_slots, p(fo), p(ss), C.byref(n)))
    t1 = time.perf_counter()
    ss0 = ss[0].copy(); nid2 = C.c_uint32(9); ps = C.c_uint32(0); th = C.c_float(0); f = C.c_float(0)
    ctx.check(L.find_pss_and_fine_timing(ctx.h, p(i_s), p(q_s), i_s.size, 4, 6, p(ss0), C.byref(nid2), C.byref(ps), C.byref(th), C.byref(f)))
    t2 = time.perf_counter()
    nid1 = C.c_uint32(0); fsi = C.c_uint32(0)
    L.find_sss(ctx.h, p(i_s), p(q_s), i_s.size, 4, 6, nid2.value, p(ss0), th.value, C.byref(nid1), C.byref(fsi))
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (nid2.value, nid1.value, fsi.value)


gpu_once()
g = [gpu_once() for _ in range(5)]
row = dict(N_slots=N_slots, samples=int(i_s.size), gpu_ms=dict(coarse=round(min(x[0] for x in g), 3), pss=round(min(x[1] for x in g), 3), sss=round(min(x[2] for x in g), 3)),
           found=dict(N_id_2=g[0][3][0], N_id_1=g[0][3][1], frame_start_idx=g[0][3][2]),
           note="host-pointer twins: every call includes the host->device copy of the I/Q window and the device->host copy of the result")
try:
    from oracle import ref as refo
    if refo.available():
        phy = refo.RefPhy(fs=30.72, N_id_cell=0xFFFF, N_ant=4, N_rb_dl=6)
        t0 = time.perf_counter(); err, fo, ss, n = phy.find_coarse_timing(i_s, q_s, N_slots); t1 = time.perf_counter()
        err, ss2, nid2, ps, th, f = phy.find_pss(i_s, q_s, ss[0]); t2 = time.perf_counter()
        phy.find_sss(i_s, q_s, nid2, ss2, th); t3 = time.perf_counter()
        row["reference_cpu_1core_ms"] = dict(coarse=round((t1 - t0) * 1e3, 2), pss=round((t2 - t1) * 1e3, 2), sss=round((t3 - t2) * 1e3, 2))
except Exception as e:      # noqa: BLE001
    row["reference_cpu_1core_ms"] = str(e)
print(json.dumps(row))
ctx.close()

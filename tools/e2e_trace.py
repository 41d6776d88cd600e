"""One traced call of lteb200_dl_receive_host on the bench workload (LTEB200_E2E_TRACE=1 prints the device-side time line); development aid."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, openlte_b200
from openlte_b200.lib import Alloc
n_sf = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = openlte_b200.load(); ctx = L.new_ctx(0)
iq, start, subfr, tb = bench.make_stream(n_sf, seed=5)
h_iq = torch.from_numpy(iq).pin_memory()
al = (Alloc * 4)(*[Alloc.make(a["mod"], a["prb"], a["tbs"], a["rnti"]) for a in bench.ALLOCS])
h_tb = torch.zeros((n_sf, 4, 6120), dtype=torch.uint8).pin_memory(); h_ok = torch.zeros((n_sf, 4), dtype=torch.int32).pin_memory()
p = lambda a: a.ctypes.data_as(C.c_void_p)
call = lambda: ctx.check(L.dl_receive_host(ctx.h, C.c_void_p(h_iq.data_ptr()), 1, iq.shape[0], 4, 100, bench.CELL, 1, p(start), p(subfr), n_sf, 2, al, 4, 1, 8,
                                           C.c_void_p(h_tb.data_ptr()), 6120, C.c_void_p(h_ok.data_ptr())))
for _ in range(3):
    call()
os.environ["LTEB200_E2E_TRACE"] = "1"
call()
del os.environ["LTEB200_E2E_TRACE"]
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    call()
torch.cuda.synchronize(); print("wall ms per call", (time.perf_counter() - t0) / 5 * 1e3, "crc ok", bool((h_ok == 1).all()))

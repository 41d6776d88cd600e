"""A few steps of the bench workload (small batch) for ncu captures of the chain kernels (development aid)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, openlte_b200
from openlte_b200.lib import Alloc
n_sf = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = openlte_b200.load(); ctx = L.new_ctx(0)
iq, start, subfr, tb = bench.make_stream(n_sf, seed=5)
d_iq = torch.from_numpy(iq).cuda()
al = (Alloc * 4)(*[Alloc.make(a["mod"], a["prb"], a["tbs"], a["rnti"]) for a in bench.ALLOCS])
d_tb = torch.zeros((n_sf, 4, 6120), dtype=torch.uint8, device="cuda"); d_ok = torch.zeros((n_sf, 4), dtype=torch.int32, device="cuda")
p = lambda a: a.ctypes.data_as(C.c_void_p)
for _ in range(3):
    ctx.check(L.dl_receive_batch(ctx.h, C.c_void_p(d_iq.data_ptr()), 1, iq.shape[0], 4, 100, bench.CELL, 1, p(start), p(subfr), n_sf, 2, al, 4, mode, 8,
                                 C.c_void_p(d_tb.data_ptr()), 6120, C.c_void_p(d_ok.data_ptr()), None))
torch.cuda.synchronize()
print("crc ok", bool(d_ok.all().item()))

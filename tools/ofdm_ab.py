"""A/B of the OFDM demodulator's sample load on the bench workload (4000 subframes, 20 MHz, int8 I/Q): sixteen 2-byte global loads per thread
(shipped) against one TMA bulk copy of the symbol's 4 KB into shared memory (LTEB200_OFDM_TMA=1).  Prints one JSON line (development aid;
output kept as profiles/r2_ofdm_tma_ab.json)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, openlte_b200
n_sf = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = openlte_b200.load(); ctx = L.new_ctx(0)
iq, start, subfr, tb = bench.make_stream(n_sf, seed=5)
d_iq = torch.from_numpy(iq).cuda()
p = lambda a: a.ctypes.data_as(C.c_void_p)
grid = [torch.zeros((n_sf, 16, 1200, 2), dtype=torch.float32, device="cuda") for _ in range(2)]
def run(g):
    ctx.check(L.ofdm_demod_batch(ctx.h, C.c_void_p(d_iq.data_ptr()), 1, iq.shape[0], 4, 100, p(start), n_sf, C.c_void_p(g.data_ptr()), None))
def t_of(g, reps=20):
    for _ in range(3): run(g)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): run(g)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
res = {}
for rep in range(2):
    os.environ.pop("LTEB200_OFDM_TMA", None)
    res.setdefault("direct_global_loads_ms", []).append(round(t_of(grid[0]), 4))
    os.environ["LTEB200_OFDM_TMA"] = "1"
    res.setdefault("tma_bulk_copy_staged_ms", []).append(round(t_of(grid[1]), 4))
os.environ.pop("LTEB200_OFDM_TMA", None)
res["grids_bit_identical"] = bool(torch.equal(grid[0], grid[1]))
res["what"] = "ofdm_demod_r16_kernel<2048, char2>, %d subframes x 16 symbols, CUDA events over 20 launches, two alternating repeats" % n_sf
print(json.dumps(res))

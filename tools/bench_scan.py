"""The scan loop at scale (LTE_fdd_dl_fs_samp_buf.cc:491-515: front end -> PDCCH blind decode -> DCI -> PDSCH of every announced allocation) on a
device-generated stream: n_sf consecutive 20 MHz subframes, each with PCFICH + 3 DCI 1A (SI-, RA-, P-RNTI) + their QPSK PDSCH + PSS/SSS/PBCH where
they belong (lteb200_dl_transmit_frames_batch), received by lteb200_dl_scan_batch in one call; beside it the reference's own per-subframe scan step
(liblte_phy_get_dl_subframe_and_ce + liblte_phy_pdcch_channel_decode + liblte_phy_pdsch_channel_decode per allocation, oracle/_ref) on host threads.
Prints one JSON line."""
import ctypes as C, json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openlte_b200
from openlte_b200.lib import Alloc, TxControl, ScanResult

n_sf = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = openlte_b200.load(); ctx = L.new_ctx(0)
n_rb, cell, cfi = 100, 17, 2
allocs = [(0xFFFF, 9, 10, 4), (0x0002, 20, 40, 9), (0xFFFE, 2, 60, 3)]            # rnti, mcs, first PRB, PRBs
tbs = [L.dci_1a_tbs(a[1]) for a in allocs]
n_alloc = len(allocs)
ca = (Alloc * n_alloc)(*[Alloc.make(1, list(range(a[2], a[2] + a[3])), t, a[0]) for a, t in zip(allocs, tbs)])
rng = np.random.default_rng(1)
gen = torch.Generator(device="cuda"); gen.manual_seed(5)
tb = torch.zeros((n_sf + 1, n_alloc, 2216), dtype=torch.uint8, device="cuda")
for a in range(n_alloc):
    tb[:, a, : tbs[a]] = torch.randint(0, 2, (n_sf + 1, tbs[a]), dtype=torch.uint8, device="cuda", generator=gen)
sn = ((np.arange(n_sf + 1)) % 10).astype(np.int32)
mib = rng.integers(0, 2, ((n_sf + 10) // 10 + 1, 24)).astype(np.uint8)
ctl = TxControl()
ctl.N_id_cell, ctl.N_ant, ctl.sync, ctl.pbch, ctl.sfn_first, ctl.pdcch, ctl.phich_res = cell, 1, 1, 1, 0, 1, 1.0
ctl.ant_gain_re[0] = 1.0
ctl.mib_bits = mib.ctypes.data_as(C.c_void_p).value
for a in range(n_alloc):
    ctl.mcs[a] = allocs[a][1]
p = lambda x: x.ctypes.data_as(C.c_void_p)
iq_f = torch.zeros(((n_sf + 1) * 30720, 2), dtype=torch.float32, device="cuda")
ctx.check(L.dl_transmit_frames_batch(ctx.h, C.c_void_p(tb.data_ptr()), 2216, ca, n_alloc, 4, n_rb, C.byref(ctl), p(sn), n_sf + 1, cfi, None, C.c_void_p(iq_f.data_ptr()), None))
torch.cuda.synchronize()
scale = 100.0 / float(iq_f.abs().max())
d_iq = (iq_f * scale + torch.randn(iq_f.shape, device="cuda", generator=gen)).round_().clamp_(-127, 127).to(torch.int8).contiguous()
del iq_f
res = (ScanResult * n_sf)()
cap = n_sf * 6 * 2216
out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
used = C.c_longlong(0)
start = (np.arange(n_sf, dtype=np.int64) * 30720)


def scan():
    ctx.check(L.dl_scan_batch(ctx.h, C.c_void_p(d_iq.data_ptr()), 1, int(d_iq.shape[0]), 4, n_rb, cell, 1, C.c_float(1.0), p(start), p(sn[:n_sf]), n_sf, 1, 8, res,
                              C.c_void_p(out.data_ptr()), cap, C.byref(used), None))


for _ in range(2):
    scan()
torch.cuda.synchronize(); t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    scan()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
bits = out.cpu().numpy(); tb_h = tb.cpu().numpy()
n_ok = n_all = 0
for b in range(n_sf):
    r = res[b]
    seen = set()
    for k in range(r.N_alloc):
        d = r.dci[k]
        a = [i for i in range(n_alloc) if allocs[i][0] == d.alloc.rnti]
        if a and r.crc_ok[k] and np.array_equal(bits[r.tb_offset[k]: r.tb_offset[k] + d.alloc.tbs], tb_h[b, a[0], : d.alloc.tbs]):
            seen.add(a[0])
    n_ok += len(seen); n_all += n_alloc
line = {"workload": "scan loop: %d consecutive 20 MHz subframes, PCFICH + 3 x DCI 1A (SI/RA/P-RNTI) + their QPSK PDSCH (tbs %s) + PSS/SSS/PBCH, int8 capture format ~30 dB" % (n_sf, tbs),
        "gpu": {"api": "lteb200_dl_scan_batch (one call; DCIs come back to the host between PDCCH and PDSCH)", "wall_ms": round(dt * 1e3, 3),
                "subframes_per_s": round(n_sf / dt, 1), "transport_blocks_recovered": "%d of %d" % (n_ok, n_all)}}
# ---- the reference's scan step on the host
try:
    from oracle import ref as refo
    if refo.available():
        iq = d_iq[: 129 * 30720].cpu().numpy()
        i_s = np.ascontiguousarray(iq[:, 0].astype(np.float32)); q_s = np.ascontiguousarray(iq[:, 1].astype(np.float32))
        n_thr = min(os.cpu_count() or 1, 64); per = 2
        okc = [0] * n_thr

        def work(t):
            phy = refo.RefPhy(fs=30.72, N_id_cell=cell, N_ant=1, N_rb_dl=n_rb); sf = refo.Subframe()
            bar.wait()
            for b in range(t * per, (t + 1) * per):
                fstart = (b * 30720 - int(sn[b]) * 30720) & 0xFFFFFFFF
                phy.get_dl_subframe_and_ce(i_s, q_s, fstart, int(sn[b]), cell, 1, sf)
                err, c, al = phy.pdcch_decode(sf, cell, 1)
                for x in al:
                    e2, bb = phy.pdsch_decode(sf, c, cell, 1, x["mod_type"], x["prb"], x["tbs"], x["rnti"], rv=x["rv_idx"])
                    okc[t] += int(e2 == 0)
        bar = threading.Barrier(n_thr + 1)
        ths = [threading.Thread(target=work, args=(t,)) for t in range(n_thr)]
        [t.start() for t in ths]; bar.wait(); t0 = time.perf_counter(); [t.join() for t in ths]; dtc = time.perf_counter() - t0
        line["reference_cpu"] = {"what": "liblte_phy_get_dl_subframe_and_ce + liblte_phy_pdcch_channel_decode + liblte_phy_pdsch_channel_decode per allocation (oracle/_ref), %d threads x %d subframes" % (n_thr, per),
                                 "subframes_per_s": round(n_thr * per / dtc, 1), "transport_blocks_decoded": int(sum(okc))}
except Exception as e:      # noqa: BLE001
    line["reference_cpu"] = {"unavailable": repr(e)}
print(json.dumps(line))

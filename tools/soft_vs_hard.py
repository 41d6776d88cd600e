"""Block error rate of the rate-0.93 "north-star load" (12 x [8 PRB, 64QAM, K = 6144] per 20 MHz subframe) against the noise level, with the
reference's hard 64QAM decisions and with the max-log soft demapper (lteb200_set_demapper); 8-iteration max-log-MAP in both cases.  Subframes 0 and 5
are left out: there the centre allocations lose resource elements to PBCH/PSS/SSS and their code rate exceeds 1 (no demapper can decode them).
Prints one JSON line per noise level; profiles/r2_soft_vs_hard_bler.jsonl is its output on the B200."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openlte_b200
from openlte_b200 import synth
from openlte_b200.lib import Alloc

L = openlte_b200.load(); ctx = L.new_ctx(0)
cell, cfi = 42, 2
allocs = [dict(mod=3, prb=list(range(8 * a, 8 * a + 8)), tbs=6120, rnti=0x300 + a) for a in range(12)]
ca = (Alloc * 12)(*[Alloc.make(a["mod"], a["prb"], a["tbs"], a["rnti"]) for a in allocs])
p = lambda a: a.ctypes.data_as(C.c_void_p)
noises = [float(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0.0, 0.6, 0.8, 1.0, 1.2, 1.4, 1.6, 1.8, 2.0]
for noise in noises:
    tot = {0: [0, 0], 1: [0, 0]}
    snr = None
    for first, seed in ((1, 3), (6, 4), (1, 5), (6, 6)):
        n_sf = 4
        tx = synth.make_subframes(n_sf, allocs, N_id_cell=cell, N_rb_dl=100, fs_idx=4, cfi=cfi, seed=seed, noise=noise, first_subfr=first)
        stream = np.concatenate([tx["samples"].reshape(-1), synth.crs_only_tail(100, cell, (first + n_sf) % 10, 4)[:8192]])
        if snr is None and noise > 0:
            clean = synth.make_subframes(n_sf, allocs, N_id_cell=cell, N_rb_dl=100, fs_idx=4, cfi=cfi, seed=seed, noise=0.0, first_subfr=first)["samples"].reshape(-1)
            # per resource element: the signal occupies 1200 of 2048 bins, the noise all of them
            snr = 10 * np.log10((np.abs(clean) ** 2).mean() / (2 * noise ** 2) * 2048 / 1200)
        iq = np.stack([stream.real, stream.imag], -1).astype(np.float32)
        d_iq = torch.from_numpy(iq).cuda()
        start = (np.arange(n_sf, dtype=np.int64) * 30720); sn = tx["subfr_nums"].astype(np.int32)
        want = np.stack(tx["tb"], 1)
        for demap in (0, 1):
            tb = torch.zeros((n_sf, 12, 6120), dtype=torch.uint8, device="cuda"); ok = torch.zeros((n_sf, 12), dtype=torch.int32, device="cuda")
            ctx.check(L.set_demapper(ctx.h, demap))
            ctx.check(L.dl_receive_batch(ctx.h, C.c_void_p(d_iq.data_ptr()), 0, iq.shape[0], 4, 100, cell, 1, p(start), p(sn), n_sf, cfi, ca, 12, 1, 8,
                                         C.c_void_p(tb.data_ptr()), 6120, C.c_void_p(ok.data_ptr()), None))
            torch.cuda.synchronize()
            L.set_demapper(ctx.h, 0)
            good = (tb.cpu().numpy() == want).all(axis=2)
            tot[demap][0] += int(good.sum()); tot[demap][1] += good.size
    print(json.dumps({"noise_per_dim": noise, "snr_per_re_dB": None if snr is None else round(float(snr), 2), "code_rate": round(6144 / (8 * 138 * 6), 3),
                      "bler_hard_64qam (reference demapper)": round(1 - tot[0][0] / tot[0][1], 4), "bler_maxlog_soft": round(1 - tot[1][0] / tot[1][1], 4),
                      "blocks": tot[0][1]}), flush=True)

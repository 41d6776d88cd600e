"""BASELINE config 5: bandwidth x modulation sweep through the whole receive chain on one GPU (development / reporting
aid; bench.py is the contract benchmark).  For every (bandwidth, modulation) the band is cut into equal allocations, each a
single-code-block transport block with the largest K for which G >= 3(K+4) in every subframe class -- the envelope in which
the reference itself decodes.  Prints one JSON line per configuration."""
import sys, os, json, ctypes as C, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openlte_b200
from openlte_b200 import synth
from openlte_b200.lib import Alloc

BW = {"1.4": (6, 0), "5": (25, 2), "10": (50, 3), "20": (100, 4)}
MODS = {"QPSK": 1, "16QAM": 2, "64QAM": 3}
QM = {1: 2, 2: 4, 3: 6}
TABLE_K = [k for lo, hi, st in ((40, 512, 8), (528, 1024, 16), (1056, 2048, 32), (2112, 6144, 64)) for k in range(lo, hi + 1, st)]


def plan(n_rb, mod, cell, cfi):
    """equal PRB groups; K = largest table size with 3(K+4) <= min_G over subframe classes, capped at 6144"""
    qm = QM[mod]
    best = None
    for n_alloc in range(1, n_rb + 1):
        if n_rb % n_alloc:
            continue
        per = n_rb // n_alloc
        groups = [list(range(a * per, (a + 1) * per)) for a in range(n_alloc)]
        gmin = min(synth.re_indices(n_rb, cell, sn, cfi, g).shape[0] * qm for g in groups for sn in (0, 5, 1))
        ks = [k for k in TABLE_K if 3 * (k + 4) <= gmin]
        if not ks:
            break
        k = ks[-1]
        if best is None or k * n_alloc > best[0]:
            best = (k * n_alloc, k, groups)
        if k == 6144:
            continue
    return best[1], best[2]


def main():
    """Single process: one GPU.  Under torchrun (BASELINE config 5: 8 x B200) every rank sweeps its own GPU with its own data -- the chain has no
    inter-GPU dependency -- and rank 0 prints whole-job subframes/s: all ranks' subframes over the slowest rank's device time."""
    argv = [a for a in sys.argv[1:] if a != "--stages"]
    want_stages = "--stages" in sys.argv          # per-stage device times (CUDA events) of the max-log-MAP chain, per configuration
    n_sf = int(argv[0]) if argv else 2000
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = openlte_b200.load(); ctx = L.new_ctx(local)
    cell, cfi = 42, 2
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    configs = [(bw, n_rb, fs_idx, mname, mod, None) for bw, (n_rb, fs_idx) in BW.items() for mname, mod in MODS.items()]
    # SURVEY 8d "north-star load": 12 x K=6144 per 20 MHz subframe (8 PRB each, code rate 0.93): punctured code words, which the
    # reference's decoder cannot decode -- GPU only, max-log-MAP is the decoder that applies
    configs.append(("20", 100, 4, "64QAM north-star load (rate 0.93; reference cannot decode this format)", 3, (6144, [list(range(8 * a, 8 * a + 8)) for a in range(12)])))
    for bw, n_rb, fs_idx, mname, mod, forced in configs:
        if True:
            K, groups = forced if forced else plan(n_rb, mod, cell, cfi)
            allocs = [dict(mod=mod, prb=g, tbs=K - 24, rnti=0x300 + i) for i, g in enumerate(groups)]
            seg = 20
            tx = synth.make_subframes(seg, allocs, N_id_cell=cell, N_rb_dl=n_rb, fs_idx=fs_idx, cfi=cfi, seed=7 + rank, first_subfr=1)
            n_s = 1920 << fs_idx
            stream = np.tile(tx["samples"].reshape(-1), -(-n_sf // seg))[: n_sf * n_s]
            stream = np.concatenate([stream, synth.crs_only_tail(n_rb, cell, (1 + n_sf) % 10, fs_idx)[: 4 * (128 << fs_idx)]])
            sc = 100.0 / max(np.abs(stream.real).max(), np.abs(stream.imag).max())
            iq = np.stack([np.rint(stream.real * sc), np.rint(stream.imag * sc)], -1).astype(np.int8)
            d_iq = torch.from_numpy(iq).cuda()
            start = (np.arange(n_sf, dtype=np.int64) * n_s); sn = ((1 + np.arange(n_sf)) % 10).astype(np.int32)
            ca = (Alloc * len(allocs))(*[Alloc.make(a["mod"], a["prb"], a["tbs"], a["rnti"]) for a in allocs])
            tb = torch.zeros((n_sf, len(allocs), K - 24), dtype=torch.uint8, device="cuda"); ok = torch.zeros((n_sf, len(allocs)), dtype=torch.int32, device="cuda")
            res = {}
            for mode, name in ((1, "maxlogmap_x8"), (0, "ref_exact")):
                run = lambda: ctx.check(L.dl_receive_batch(ctx.h, C.c_void_p(d_iq.data_ptr()), 1, iq.shape[0], fs_idx, n_rb, cell, 1, p(start), p(sn), n_sf, cfi,
                                                           ca, len(allocs), mode, 8, C.c_void_p(tb.data_ptr()), K - 24, C.c_void_p(ok.data_ptr()), None))
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    run()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                if world > 1:
                    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
                want = np.stack(tx["tb"], 1)[np.arange(n_sf) % seg]
                okf = torch.tensor([float(ok.float().mean().item()), float(np.array_equal(tb.cpu().numpy(), want))], device="cuda")
                if world > 1:
                    dist.all_reduce(okf, op=dist.ReduceOp.MIN)
                res[name] = dict(ms_per_step=round(ms, 3), subframes_per_s=round(world * n_sf / ms * 1e3, 1), crc_pass=float(okf[0].item()),
                                 bits_ok=bool(okf[1].item() == 1.0))
            if want_stages and rank == 0:
                nsc = 12 * n_rb
                grid = torch.zeros((n_sf, 16, nsc, 2), dtype=torch.float32, device="cuda"); ce = torch.zeros((n_sf, 1, 14, nsc, 2), dtype=torch.float32, device="cuda")
                vp = lambda t_: C.c_void_p(t_.data_ptr())

                def t_of(fn, reps=3):
                    fn(); torch.cuda.synchronize()
                    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(reps):
                        fn()
                    b.record(); torch.cuda.synchronize()
                    return round(a.elapsed_time(b) / reps, 4)
                st = {}
                st["ofdm_demod"] = t_of(lambda: ctx.check(L.ofdm_demod_batch(ctx.h, vp(d_iq), 1, iq.shape[0], fs_idx, n_rb, p(start), n_sf, vp(grid), None)))
                st["chan_est"] = t_of(lambda: ctx.check(L.chan_est_batch(ctx.h, vp(grid), p(sn), n_rb, cell, 1, n_sf, vp(ce), None)))
                stride = int(L.q8_stride(K))
                if stride:
                    # (the stage entry points take at most 8 allocations per call)
                    groups8 = [list(range(g, min(g + 8, len(allocs)))) for g in range(0, len(allocs), 8)]
                    cas = [(Alloc * len(g))(*[ca[i] for i in g]) for g in groups8]
                    q8s = [torch.zeros((n_sf * len(g), stride), dtype=torch.int8, device="cuda") for g in groups8]
                    tbs_ = [torch.zeros((n_sf, len(g), K - 24), dtype=torch.uint8, device="cuda") for g in groups8]
                    oks_ = [torch.zeros((n_sf, len(g)), dtype=torch.int32, device="cuda") for g in groups8]
                    st["pdsch_llr_q8"] = t_of(lambda: [ctx.check(L.pdsch_llr_q8_batch(ctx.h, vp(grid), vp(ce), n_rb, cell, 1, p(sn), n_sf, cfi, cas[i], len(g), vp(q8s[i]), None))
                                                       for i, g in enumerate(groups8)])
                    st["turbo_mlm2"] = t_of(lambda: [ctx.check(L.turbo_decode_q8_batch(ctx.h, vp(q8s[i]), K, n_sf * len(g), 8, cas[i], len(g), n_sf, vp(tbs_[i]), K - 24, vp(oks_[i]), None))
                                                     for i, g in enumerate(groups8)])
                else:
                    st["pdsch_decode (fused window-per-thread kernel: K is not a multiple of 32)"] = t_of(lambda: ctx.check(
                        L.pdsch_decode_batch(ctx.h, vp(grid), vp(ce), n_rb, cell, 1, p(sn), n_sf, cfi, ca, len(allocs), 1, 8, vp(tb), K - 24, vp(ok), None)))
                st["sum"] = round(sum(st.values()), 4)
                res["stages_ms_maxlogmap"] = st
            if rank == 0:
                print(json.dumps(dict(bandwidth_MHz=bw, N_rb_dl=n_rb, modulation=mname, n_alloc=len(allocs), prb_per_alloc=len(groups[0]), K=K, n_sf_per_gpu=n_sf, n_gpus=world,
                                      **res)), flush=True)


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.destroy_process_group()

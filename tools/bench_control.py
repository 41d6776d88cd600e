"""Throughput of the control-channel kernels on one GPU, beside the reference's CPU functions.
Input: the reference-generated fixtures under tests/golden/ replicated into a batch (a 20 MHz control region with PCFICH and three
DCI 1A; a 20 MHz PBCH subframe).  Prints one JSON line per kernel.  Development / reporting aid (bench.py is the contract benchmark)."""
import sys, os, json, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import openlte_b200
from openlte_b200.lib import BchResult, PdcchRaw

GOLD = os.path.join(ROOT, "tests", "golden")


def planes(a):
    out = np.zeros(a.shape[:-2] + (16, 1200), np.float32); out[..., :a.shape[-2], :a.shape[-1]] = a
    return out


def t_gpu(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    L = openlte_b200.load(); ctx = L.new_ctx(0)
    vp = lambda t: C.c_void_p(t.data_ptr())
    z = np.load(os.path.join(GOLD, "control_pdcch_ref.npz"))
    nrb, cell, cfi, subfr = [int(x) for x in z["cfg_1"][:4]]
    grid1 = torch.from_numpy(np.stack([z["rx_re_1"], z["rx_im_1"]], -1).astype(np.float32)).cuda()
    ce1 = torch.from_numpy(np.stack([z["ce_re_1"][:14], z["ce_im_1"][:14]], -1).astype(np.float32)).cuda()
    grid = grid1.unsqueeze(0).repeat(B, 1, 1, 1).contiguous(); ce = ce1.unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
    out = torch.zeros(B * C.sizeof(PdcchRaw), dtype=torch.uint8, device="cuda")
    sn = np.full(B, subfr, np.int32)
    ms = t_gpu(lambda: ctx.check(L.pdcch_decode_batch(ctx.h, vp(grid), vp(ce), nrb, cell, 1, C.c_float(1.0), sn.ctypes.data_as(C.c_void_p), B, vp(out), None)))
    raw = (PdcchRaw * B).from_buffer_copy(out.cpu().numpy().tobytes())
    n_found = sum(raw[0].found)
    cpu_ms = None
    try:
        from oracle import ref as refo
        if refo.available():
            phy = refo.RefPhy(fs=30.72, N_id_cell=cell, N_ant=1, N_rb_dl=nrb); sf = refo.Subframe(); sf.clear(); sf.set_num(subfr)
            sf.rx_symb_re[:, :12 * nrb] = z["rx_re_1"]; sf.rx_symb_im[:, :12 * nrb] = z["rx_im_1"]
            sf.rx_ce_re[0, :, :12 * nrb] = z["ce_re_1"]; sf.rx_ce_im[0, :, :12 * nrb] = z["ce_im_1"]
            fd = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 1)      # the reference printf()s from inside the decoder
            try:
                t0 = time.perf_counter()
                for _ in range(50):
                    phy.pdcch_decode2(sf, cell, 1)
                cpu_ms = (time.perf_counter() - t0) / 50 * 1e3
            finally:
                os.dup2(fd, 1)
    except Exception as e:      # noqa: BLE001
        cpu_ms = None
    print(json.dumps(dict(kernel="pdcch_decode_kernel (PCFICH + 12 blind decodes per subframe)", N_rb_dl=nrb, subframes=B, ms_per_batch=round(ms, 4),
                          subframes_per_s=round(B / ms * 1e3, 1), blind_decodes_per_s=round(12 * B / ms * 1e3, 1), dcis_found_per_subframe=int(n_found),
                          cpu_reference_ms_per_subframe_1core=None if cpu_ms is None else round(cpu_ms, 4),
                          speedup_vs_1core=None if cpu_ms is None else round(cpu_ms * B / ms, 1))), flush=True)

    z = np.load(os.path.join(GOLD, "control_bch_ref.npz"))
    nrb, cell = int(z["cfg_0"][0]), int(z["cfg_0"][1])
    g1 = torch.from_numpy(np.stack([z["rx_re_0"], z["rx_im_0"]], -1).astype(np.float32)).cuda()
    c1 = torch.from_numpy(np.stack([z["ce_re_0"][:, :14], z["ce_im_0"][:, :14]], -1).astype(np.float32)).cuda()
    Bb = min(B, 1000)
    grid = g1.unsqueeze(0).repeat(Bb, 1, 1, 1).contiguous(); ce = c1.unsqueeze(0).repeat(Bb, 1, 1, 1, 1).contiguous()
    out = torch.zeros(Bb * C.sizeof(BchResult), dtype=torch.uint8, device="cuda")
    ms = t_gpu(lambda: ctx.check(L.bch_decode_batch(ctx.h, vp(grid), vp(ce), nrb, cell, Bb, vp(out), None)))
    res = (BchResult * Bb).from_buffer_copy(out.cpu().numpy().tobytes())
    cpu_ms = None
    try:
        from oracle import ref as refo
        if refo.available():
            phy = refo.RefPhy(fs=30.72, N_id_cell=cell, N_ant=4, N_rb_dl=nrb); sf = refo.Subframe(); sf.clear(); sf.set_num(0)
            sf.rx_symb_re[:, :12 * nrb] = z["rx_re_0"]; sf.rx_symb_im[:, :12 * nrb] = z["rx_im_0"]
            sf.rx_ce_re[:, :, :12 * nrb] = z["ce_re_0"]; sf.rx_ce_im[:, :, :12 * nrb] = z["ce_im_0"]
            t0 = time.perf_counter()
            for _ in range(50):
                phy.bch_decode(sf, cell)
            cpu_ms = (time.perf_counter() - t0) / 50 * 1e3
    except Exception:           # noqa: BLE001
        cpu_ms = None
    print(json.dumps(dict(kernel="bch_decode_kernel (3 antenna hypotheses x 4 offsets per frame)", N_rb_dl=nrb, frames=Bb, ms_per_batch=round(ms, 4),
                          frames_per_s=round(Bb / ms * 1e3, 1), all_ok=all(r.ok for r in res),
                          cpu_reference_ms_per_frame_1core=None if cpu_ms is None else round(cpu_ms, 4),
                          note="the reference stops at the first hypothesis that passes (here the first); the GPU always evaluates all 12")), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

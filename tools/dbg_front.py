import sys, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import openlte_b200
from oracle import port
L = openlte_b200.load(); ctx = L.new_ctx(0)
z = np.load('/root/repo/tests/golden/loop_64qam_c0_sf1.npz'); case = {k: z[k] for k in z.files}
iq = np.stack([case["i"], case["q"]], axis=-1).astype(np.float32)
d_iq = torch.from_numpy(iq).cuda()
nsc = 1200
grid = torch.zeros((1, 16, nsc, 2), dtype=torch.float32, device="cuda")
start = np.array([int(case["frame_start"]) + int(case["subfr"]) * int(case["n_sf"])], np.int64)
p = lambda a: a.ctypes.data_as(C.c_void_p)
rc = L.ofdm_demod_batch(ctx.h, C.c_void_p(d_iq.data_ptr()), 0, iq.shape[0], 4, 100, p(start), 1, C.c_void_p(grid.data_ptr()), None)
torch.cuda.synchronize(); print('rc', rc, L.last_error(ctx.h))
g = grid.cpu().numpy()[0]; g = g[...,0] + 1j*g[...,1]
ref = case["ref_rx_re"] + 1j*case["ref_rx_im"]
print('grid max', np.abs(ref).max(), 'err', np.abs(g-ref).max())
for s in range(16): print(s, np.abs(g[s]-ref[s]).max(), np.abs(g[s]).max())
ce = torch.zeros((1, 1, 14, nsc, 2), dtype=torch.float32, device="cuda")
sn = np.array([int(case["subfr"])], np.int32)
rc = L.chan_est_batch(ctx.h, C.c_void_p(grid.data_ptr()), p(sn), 100, int(case["N_id_cell"]), 1, 1, C.c_void_p(ce.data_ptr()), None)
torch.cuda.synchronize(); print('rc', rc)
c = ce.cpu().numpy()[0,0]; c = c[...,0]+1j*c[...,1]
rc_ = case["ref_ce_re"][0] + 1j*case["ref_ce_im"][0]
for s in range(14):
    e = np.abs(c[s]-rc_[s]); print(s, 'maxerr', e.max(), 'first bad sc', int(np.argmax(e > 1e-2)) if (e>1e-2).any() else -1, 'n bad', int((e>1e-2).sum()))
y = g[4, 3::6][:5]
print('gpu ce[4] at crs tones', c[4, 3::6][:5])
print('ref ce[4] at crs tones', rc_[4, 3::6][:5])
for ns in range(20):
    for l in (0,1,4):
        re, im = port.crs(ns, l, 0); cr = (re + 1j*im)[10:15]
        v = y*np.conj(cr)
        if np.abs(v - c[4, 3::6][:5]).max() < 1: print('GPU used crs ns', ns, 'l', l)
        if np.abs(v - rc_[4, 3::6][:5]).max() < 1: print('REF used crs ns', ns, 'l', l)

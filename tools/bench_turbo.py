"""Quick device timing of the turbo kernels (development aid; bench.py is the contract benchmark)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openlte_b200
from oracle import port

from openlte_b200.lib import Lib
L = Lib(os.environ['LTEB200_SO']) if os.environ.get('LTEB200_SO') else openlte_b200.load(); ctx = L.new_ctx(0)
K = 6144
rng = np.random.default_rng(0)
blocks = []
for s in (0.0, 0.3, 0.5, 0.8):
    c = rng.integers(0, 2, K).astype(np.uint8); d = port.turbo_encode(c)
    x = (1.0 - 2.0 * d.astype(np.float32)).reshape(3, K + 4).T.reshape(-1) + s * rng.standard_normal(3 * (K + 4))
    blocks.append(x.astype(np.float32))
for n in (1024, 4096, 16384):
    xs = torch.from_numpy(np.stack([blocks[i % 4] for i in range(n)])).cuda()
    out = torch.empty((n, K), dtype=torch.uint8, device="cuda")
    for mode, it in ((0, 0), (1, 8)):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(2):
            ctx.check(L.turbo_decode_batch(ctx.h, C.c_void_p(xs.data_ptr()), K, n, mode, it, C.c_void_p(out.data_ptr()), st))
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ctx.check(L.turbo_decode_batch(ctx.h, C.c_void_p(xs.data_ptr()), K, n, mode, it, C.c_void_p(out.data_ptr()), st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"n_cb={n} mode={'A(ref-exact)' if mode == 0 else 'B(maxlogmap x8)'}: {ms:.3f} ms  -> {n / ms * 1e3:.0f} CB/s", flush=True)

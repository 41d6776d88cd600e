import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch, openlte_b200
from oracle import port
L = openlte_b200.load(); ctx = L.new_ctx(0)
for K in (2112, 3072, 3136, 4288, 6144):
    rng = np.random.default_rng(0); c = rng.integers(0, 2, K).astype(np.uint8); d = port.turbo_encode(c)
    x = ((1.0 - 2.0 * d.astype(np.float32)).reshape(3, K + 4).T.reshape(-1) + 0.5 * rng.standard_normal(3 * (K + 4))).astype(np.float32)
    n = 8192
    xs = torch.from_numpy(np.tile(x, (n, 1))).cuda(); out = torch.empty((n, K), dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    run = lambda: ctx.check(L.turbo_decode_batch(ctx.h, C.c_void_p(xs.data_ptr()), K, n, 1, 8, C.c_void_p(out.data_ptr()), st))
    run(); run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"K={K}: {ms:.3f} ms per {n} blocks, {ms / n / K * 1e6:.4f} ns per bit")

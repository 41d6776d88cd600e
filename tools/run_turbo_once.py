"""One launch of each turbo decoder on a K = 6144 batch, for ncu captures (development aid)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openlte_b200
from oracle import port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
L = openlte_b200.load(); ctx = L.new_ctx(0)
K = 6144
rng = np.random.default_rng(0)
blocks = []
for s in (0.0, 0.3, 0.5, 0.8):
    c = rng.integers(0, 2, K).astype(np.uint8); d = port.turbo_encode(c)
    x = (1.0 - 2.0 * d.astype(np.float32)).reshape(3, K + 4).T.reshape(-1) + s * rng.standard_normal(3 * (K + 4))
    blocks.append(x.astype(np.float32))
xs = torch.from_numpy(np.stack([blocks[i % 4] for i in range(n)])).cuda()
out = torch.empty((n, K), dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode in modes:
    for _ in range(2):
        ctx.check(L.turbo_decode_batch(ctx.h, C.c_void_p(xs.data_ptr()), K, n, mode, 8, C.c_void_p(out.data_ptr()), st))
    torch.cuda.synchronize()
print("done")

"""BASELINE config 3: 1024 code blocks of K = 6144, bits from a seeded generator, encoded with the reference-compatible encoder (so the
wrapped interleaver of SURVEY H2 is baked in), BPSK +-1 with AWGN at several sigma.  Both decoders on one GPU: mode A (the reference's
turbo_decode, bit-exact -- checked against the CPU oracle on a sample of blocks here and on the whole batch in tests/test_turbo_gpu.py)
and mode B (8-iteration max-log-MAP).  Prints one JSON line per noise point: bit/block error rates against the transmitted bits and the
device time per batch."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openlte_b200
from oracle import port

K, N = 6144, 1024
L = openlte_b200.load(); ctx = L.new_ctx(0)
bits = np.stack([np.random.default_rng(1234 + b).integers(0, 2, K).astype(np.uint8) for b in range(N)])
code = np.stack([port.turbo_encode(bits[b]) for b in range(N)])                      # branch-major [3][K+4]
tx = (1.0 - 2.0 * code.astype(np.float32)).reshape(N, 3, K + 4).transpose(0, 2, 1).reshape(N, -1)   # d[3i+x]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for sigma in (0.0, 0.3, 0.5, 0.7, 0.9, 1.0):
    rng = np.random.default_rng(int(sigma * 100) + 7)
    x = (tx + sigma * rng.standard_normal(tx.shape)).astype(np.float32)
    d_x = torch.from_numpy(x).cuda()
    row = dict(sigma=sigma, EsN0_dB=None if sigma == 0 else round(-20 * np.log10(sigma), 2), n_blocks=N, K=K)
    for mode, name in ((0, "reference_exact"), (1, "maxlogmap_x8")):
        out = torch.empty((N, K), dtype=torch.uint8, device="cuda")
        run = lambda: ctx.check(L.turbo_decode_batch(ctx.h, C.c_void_p(d_x.data_ptr()), K, N, mode, 8, C.c_void_p(out.data_ptr()), st))
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record(); torch.cuda.synchronize()
        got = out.cpu().numpy()
        err = got != bits
        row[name] = dict(ber=float(err.mean()), bler=float(err.any(1).mean()), ms_per_batch=round(e0.elapsed_time(e1) / 5, 4))
        if mode == 0:          # the oracle restatement of the reference's decoder on a sample of the batch (each call modifies its input)
            same = all(np.array_equal(got[b], port.turbo_decode_ref(x[b].copy())) for b in range(0, N, 64))
            row[name]["bit_exact_with_cpu_oracle_on_sample"] = bool(same)
    print(json.dumps(row), flush=True)
ctx.close()

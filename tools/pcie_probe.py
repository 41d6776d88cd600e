"""Host<->device copy rates of this box for the e2e pipeline's transfer sizes (246 MB in, 98 MB out per step): alone, concurrent,
and in 16 chunks.  Development aid."""
import torch, time
h_in = torch.empty(245776384, dtype=torch.uint8).pin_memory(); d_in = torch.empty_like(h_in, device="cuda")
h_out = torch.empty(97984000, dtype=torch.uint8).pin_memory(); d_out = torch.empty_like(h_out, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


def chunked(n=16):
    a = h_in.numel() // n; b = h_out.numel() // n
    for c in range(n):
        with torch.cuda.stream(s1):
            d_in[c * a:(c + 1) * a].copy_(h_in[c * a:(c + 1) * a], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[c * b:(c + 1) * b].copy_(d_out[c * b:(c + 1) * b], non_blocking=True)


t = timed(h2d); print("H2D 246 MB alone   %.2f ms  %.1f GB/s" % (t, 0.2458 / t * 1e3))
t = timed(d2h); print("D2H  98 MB alone   %.2f ms  %.1f GB/s" % (t, 0.0980 / t * 1e3))
t = timed(both); print("both concurrently  %.2f ms" % t)
t = timed(chunked); print("both, 16 chunks    %.2f ms" % t)

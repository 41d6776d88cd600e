"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def noisy_codeword(port, K, sigma, rng, sentinel_every=0):
    """Random block -> reference-compatible turbo code word -> BPSK + AWGN in the decoder's d[3i+x] layout."""
    c = rng.integers(0, 2, K).astype(np.uint8)
    d = port.turbo_encode(c)                       # branch-major
    D = K + 4
    x = (1.0 - 2.0 * d.astype(np.float32)).reshape(3, D).T.reshape(-1).copy()
    if sigma > 0:
        x = (x + sigma * rng.standard_normal(x.size)).astype(np.float32)
    if sentinel_every:
        x[1::sentinel_every] = 10000.0
    return c, x.astype(np.float32)


def dptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import ctypes
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

"""LTEB200_QPP_SPEC: the turbo code's internal interleaver as 3GPP TS 36.212 5.1.3.2.3 defines it, next to the reference's uint32-wrapping
one (SURVEY 8f rank 3 names a spec-correct mode as the natural extension of the generator).  No reference counterpart, so the
pins are: the permutation property and an independent pure-Python encoder (CPU), the oracle port switched to the same interleaver (bit-exact
decoder twins), and transmit -> receive round trips (GPU)."""
import ctypes as C

import numpy as np
import pytest

from util import dptr, stream_ptr


def py_qpp(K, f1, f2):
    return np.array([(f1 * i + f2 * i * i) % K for i in range(K)], dtype=np.int64)      # Python integers: no wrap


def py_rsc(bits):
    """One constituent encoder of 36.212 5.1.3.2.1: g0 = 1 + D^2 + D^3 (feedback), g1 = 1 + D + D^3; returns parity z[K+3] and the three
    termination systematic bits x[K..K+2] (the switch in the lower position: the register is fed with its own feedback)."""
    s1 = s2 = s3 = 0
    z = []
    for b in bits:
        a = int(b) ^ s2 ^ s3
        z.append(a ^ s1 ^ s3)
        s1, s2, s3 = a, s1, s2
    xt = []
    for _ in range(3):
        f = s2 ^ s3
        xt.append(f)                # x_k = feedback, so the register input a = 0
        z.append(s1 ^ s3)
        s1, s2, s3 = 0, s1, s2
    return np.array(z, np.uint8), np.array(xt, np.uint8)


def py_turbo_encode(c, pi):
    """36.212 5.1.3.2: d0/d1/d2 with the trellis termination multiplexing of 5.1.3.2.2, branch-major [3][K+4]."""
    K = c.size
    z, x = py_rsc(c)
    zp, xp = py_rsc(c[pi])
    d = np.zeros((3, K + 4), np.uint8)
    d[0, :K] = c; d[1, :K] = z[:K]; d[2, :K] = zp[:K]
    d[0, K:] = [x[0], z[K + 1], xp[0], zp[K + 1]]
    d[1, K:] = [z[K], x[2], zp[K], xp[2]]
    d[2, K:] = [x[1], z[K + 2], xp[1], zp[K + 2]]
    return d.reshape(-1)


def test_spec_permutation_and_independent_encoder(port):
    from oracle import port as P
    ks = [k for lo, hi, st in ((40, 512, 8), (528, 1024, 16), (1056, 2048, 32), (2112, 6144, 64)) for k in range(lo, hi + 1, st)]
    wrapped = []
    try:
        for K in ks:
            f1, f2 = P.qpp_params(K)
            want = py_qpp(K, f1, f2)
            assert np.array_equal(np.sort(want), np.arange(K)), f"K={K}: the 36.212 polynomial must be a permutation"
            P.set_qpp_spec(True); spec = P.qpp_pi(K)
            P.set_qpp_spec(False); ref = P.qpp_pi(K)
            assert np.array_equal(spec, want)
            if not np.array_equal(ref, want): wrapped.append(K)
        # the reference's uint32 arithmetic departs from the specification for the larger block sizes only
        assert wrapped and min(wrapped) > 2048 and 6144 in wrapped
        rng = np.random.default_rng(3)
        for K in (40, 1024, 6144):
            c = rng.integers(0, 2, K).astype(np.uint8)
            f1, f2 = P.qpp_params(K)
            P.set_qpp_spec(True)
            assert np.array_equal(P.turbo_encode(c), py_turbo_encode(c, py_qpp(K, f1, f2))), f"K={K}: port encoder (spec interleaver) vs the pure-Python one"
            P.set_qpp_spec(False)
            if K not in wrapped:
                assert np.array_equal(P.turbo_encode(c), py_turbo_encode(c, py_qpp(K, f1, f2))), "where nothing wraps the reference's code IS the 36.212 code"
    finally:
        P.set_qpp_spec(False)


@pytest.mark.gpu
def test_spec_mode_decoders_match_their_twins(ctx, port):
    import torch
    from util import noisy_codeword
    lib = ctx.lib
    rng = np.random.default_rng(77)
    try:
        port.set_qpp_spec(True)
        ctx.check(lib.set_interleaver(ctx.h, 1))
        for K in (40, 1024, 3584, 6144):
            xs, cs = [], []
            for sigma in (0.0, 0.4, 0.7):
                c, x = noisy_codeword(port, K, sigma, rng)
                xs.append(x); cs.append(c)
            xs = np.stack(xs)
            d = torch.from_numpy(xs).cuda()
            out = torch.empty((xs.shape[0], K), dtype=torch.uint8, device="cuda")
            for mode, twin in ((0, lambda v: port.turbo_decode_ref(v)), (1, lambda v: port.turbo_decode_mlm(v, 8))):
                ctx.check(lib.turbo_decode_batch(ctx.h, dptr(d), K, xs.shape[0], mode, 8, dptr(out), stream_ptr()))
                torch.cuda.synchronize()
                got = out.cpu().numpy()
                for b in range(xs.shape[0]):
                    assert np.array_equal(got[b], twin(xs[b])), f"K={K} mode={mode} block {b}"
                assert np.array_equal(got[0], cs[0]) and (mode == 0 or np.array_equal(got[1], cs[1]))
        # the option matters: a 36.212 code word of a size where the reference wraps does not decode with the reference's interleaver
        K = 6144
        c, x = noisy_codeword(port, K, 0.0, rng)
        ctx.check(lib.set_interleaver(ctx.h, 0))
        d = torch.from_numpy(x[None]).cuda(); out = torch.empty((1, K), dtype=torch.uint8, device="cuda")
        ctx.check(lib.turbo_decode_batch(ctx.h, dptr(d), K, 1, 1, 8, dptr(out), stream_ptr()))
        torch.cuda.synchronize()
        assert not np.array_equal(out.cpu().numpy()[0], c)
    finally:
        port.set_qpp_spec(False)
        lib.set_interleaver(ctx.h, 0)


@pytest.mark.gpu
def test_spec_mode_transmit_receive_round_trip(ctx):
    """Device transmitter and receive chain both in spec mode: transport blocks come back, for both decoders; the two modes produce different
    signals for K = 6144 and each decodes only its own."""
    import torch
    from openlte_b200.lib import Alloc
    lib = ctx.lib
    cell, cfi, n_sf, tbs = 12, 2, 4, 6120
    allocs = [Alloc.make(3, list(range(25 * a, 25 * a + 25)), tbs, 0x300 + a) for a in range(2)]
    ca = (Alloc * 2)(*allocs)
    rng = np.random.default_rng(9)
    bits = torch.from_numpy(rng.integers(0, 2, (n_sf, 2, tbs)).astype(np.uint8)).cuda()
    sn = ((1 + np.arange(n_sf + 1)) % 10).astype(np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n_samps = (n_sf + 1) * 30720
    start = (np.arange(n_sf, dtype=np.int64) * 30720)

    def transmit():
        # n_sf subframes with data + one more (its first symbols are the estimator's look-ahead); unit-gain float I/Q
        tb = torch.zeros((n_sf + 1, 2, tbs), dtype=torch.uint8, device="cuda"); tb[:n_sf] = bits
        iq = torch.zeros((n_samps, 2), dtype=torch.float32, device="cuda")
        ctx.check(lib.dl_transmit_batch(ctx.h, dptr(tb), tbs, ca, 2, 4, 100, cell, 1, p(sn), n_sf + 1, cfi, None, dptr(iq), stream_ptr()))
        torch.cuda.synchronize()
        return iq

    def receive(iq, mode):
        tb = torch.zeros((n_sf, 2, tbs), dtype=torch.uint8, device="cuda"); ok = torch.zeros((n_sf, 2), dtype=torch.int32, device="cuda")
        ctx.check(lib.dl_receive_batch(ctx.h, dptr(iq), 0, n_samps, 4, 100, cell, 1, p(start), p(sn[:n_sf]), n_sf, cfi, ca, 2, mode, 8, dptr(tb), tbs, dptr(ok), stream_ptr()))
        torch.cuda.synchronize()
        return ok.cpu().numpy(), tb

    try:
        iq_ref = transmit()
        ctx.check(lib.set_interleaver(ctx.h, 1))
        iq_spec = transmit()
        assert not torch.equal(iq_ref, iq_spec)
        for mode in (1, 0):
            ok, tb = receive(iq_spec, mode)
            assert ok.all() and torch.equal(tb, bits), f"spec signal, spec receiver, decoder {mode}"
        ok, _ = receive(iq_ref, 1)
        assert not ok.any(), "a signal of the reference's transmitter does not decode with the 36.212 interleaver at K = 6144"
        ctx.check(lib.set_interleaver(ctx.h, 0))
        ok, tb = receive(iq_ref, 1)
        assert ok.all() and torch.equal(tb, bits)
    finally:
        lib.set_interleaver(ctx.h, 0)

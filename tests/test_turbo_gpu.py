"""Parity of the CUDA turbo decoders against the oracle, called through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from util import noisy_codeword, dptr, stream_ptr

pytestmark = pytest.mark.gpu


def gpu_decode(ctx, x_batch, K, mode, n_iter=8):
    import torch
    n = x_batch.shape[0]
    d = torch.from_numpy(np.ascontiguousarray(x_batch)).cuda()
    out = torch.empty((n, K), dtype=torch.uint8, device="cuda")
    ctx.check(ctx.lib.turbo_decode_batch(ctx.h, dptr(d), K, n, mode, n_iter, dptr(out), stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), torch.from_numpy(x_batch)), "batched API must not modify its input"
    return out.cpu().numpy()


@pytest.mark.parametrize("K", [40, 48, 512, 1024, 3584, 5248, 6016, 6144])
def test_mode_a_bit_exact(ctx, port, refo, K):
    rng = np.random.default_rng(100 + K)
    xs, cs = [], []
    for sigma in [0.0, 0.3, 0.5, 0.8, 1.2]:
        for rep in range(3):
            c, x = noisy_codeword(port, K, sigma, rng, sentinel_every=(21 if rep == 0 else 0))
            xs.append(x); cs.append(c)
    xs = np.stack(xs)
    got = gpu_decode(ctx, xs, K, 0)
    for b in range(xs.shape[0]):
        want = port.turbo_decode_ref(xs[b])
        assert np.array_equal(got[b], want), f"K={K} block {b}: {(got[b] != want).sum()} bits differ from the oracle port"
    if refo is not None:
        phy = refo.RefPhy()
        for b in range(0, xs.shape[0], 4):
            assert np.array_equal(got[b], phy.turbo_decode(xs[b])), f"K={K} block {b} differs from the reference itself"
    # noise-free, unpunctured blocks decode to the transmitted bits (self-consistency of the reference codec,
    # BASELINE.md P3); block 0 carries sentinels, i.e. is punctured, where the reference itself fails
    for b in (1, 2):
        assert np.array_equal(got[b], cs[b])


@pytest.mark.parametrize("K", [40, 104, 512, 1024, 3520, 6016, 6144])
def test_mode_b_matches_cpu_twin(ctx, port, K):
    rng = np.random.default_rng(200 + K)
    xs, cs, sig = [], [], []
    for sigma in [0.0, 0.5, 0.8, 1.0]:
        for rep in range(2):
            c, x = noisy_codeword(port, K, sigma, rng, sentinel_every=(33 if rep else 0))
            xs.append(x); cs.append(c); sig.append(sigma)
    xs = np.stack(xs)
    for n_iter in (1, 8):
        got = gpu_decode(ctx, xs, K, 1, n_iter)
        for b in range(xs.shape[0]):
            want = port.turbo_decode_mlm(xs[b], n_iter)
            assert np.array_equal(got[b], want), f"K={K} it={n_iter} block {b}: {(got[b] != want).sum()} bits differ from the CPU twin"
    got = gpu_decode(ctx, xs, K, 1, 8)
    for b in range(xs.shape[0]):
        if sig[b] == 0.0 or (K >= 512 and K < 3584 and sig[b] <= 0.8):
            assert np.array_equal(got[b], cs[b]), f"K={K} sigma={sig[b]}: max-log-MAP failed to recover the block"


def test_mode_b_equals_reference_where_both_decode(ctx, port):
    """North-star contract: decoded bits equal the reference's wherever both decode correctly."""
    rng = np.random.default_rng(7)
    for K in (1024, 6144):
        xs, cs = zip(*[noisy_codeword(port, K, 0.0 if i < 4 else 0.25, rng)[::-1] for i in range(8)])
        xs = np.stack(xs)
        a = gpu_decode(ctx, xs, K, 0)
        b = gpu_decode(ctx, xs, K, 1, 8)
        for i in range(8):
            if np.array_equal(a[i], cs[i]):
                assert np.array_equal(b[i], a[i])


def test_config3_full_batch(ctx, port):
    """BASELINE config 3: 1024 code blocks of K = 6144, mode A bit-exact at three noise points."""
    K, n = 6144, 1024
    rng = np.random.default_rng(1234)
    base = [noisy_codeword(port, K, s, rng)[1] for s in (0.0, 0.3, 0.5) for _ in range(8)]
    xs = np.stack([base[i % len(base)] * (1.0 if i < len(base) else (1.0 + 0.001 * (i % 7))) for i in range(n)]).astype(np.float32)
    got = gpu_decode(ctx, xs, K, 0)
    idx = list(range(0, 24)) + list(rng.integers(24, n, 40))
    for b in idx:
        assert np.array_equal(got[b], port.turbo_decode_ref(xs[b])), f"block {b}"
    # identical inputs decode identically wherever they sit in the batch (no cross-block state)
    scale_free = [i for i in range(24, n) if (i % 7) == 0]
    for i in scale_free[:50]:
        assert np.array_equal(got[i], got[i % len(base)])


def test_host_twin_modifies_input_like_reference(ctx, port):
    K = 512
    rng = np.random.default_rng(5)
    c, x = noisy_codeword(port, K, 0.4, rng, sentinel_every=9)
    xin = x.copy()
    out = np.zeros(K, np.uint8)
    rc = ctx.lib.turbo_decode(ctx.h, xin.ctypes.data_as(C.c_void_p), xin.size, 0, out.ctypes.data_as(C.c_void_p), 0, 0)
    ctx.check(rc)
    assert not (xin == 10000.0).any() and (xin[x == 10000.0] == 0).all()      # liblte_phy.cc:10636-10642
    assert np.array_equal(out, port.turbo_decode_ref(x))


def test_bad_arguments(ctx):
    import torch
    d = torch.zeros(3 * 47, device="cuda"); o = torch.zeros(64, dtype=torch.uint8, device="cuda")
    assert ctx.lib.turbo_decode_batch(ctx.h, dptr(d), 43, 1, 0, 0, dptr(o), None) == -1   # 43 is not a block size
    assert ctx.lib.turbo_decode_batch(ctx.h, None, 40, 1, 0, 0, dptr(o), None) == -1


def test_every_legal_block_size(ctx, port):
    """All 188 sizes of 36.212 table 5.1.3-3 (including the 20 whose interleaver the reference's uint32 arithmetic breaks):
    one noisy and one clean block each, both decoders, against the oracle."""
    import re, os
    rows = re.findall(r"\{\s*(\d+),\s*(\d+),\s*(\d+)\}", open(os.path.join(os.path.dirname(__file__), "..", "openlte_b200", "csrc", "qpp_table.h")).read())
    Ks = [int(r[0]) for r in rows]
    assert len(Ks) == 188
    rng = np.random.default_rng(2024)
    for K in Ks:
        blocks = [noisy_codeword(port, K, s, rng)[1] for s in (0.0, 0.6)]
        xs = np.stack(blocks)
        a = gpu_decode(ctx, xs, K, 0)
        b = gpu_decode(ctx, xs, K, 1, 4)
        for i in range(2):
            assert np.array_equal(a[i], port.turbo_decode_ref(xs[i])), f"mode A, K={K}"
            assert np.array_equal(b[i], port.turbo_decode_mlm(xs[i], 4)), f"mode B, K={K}"


def test_empty_and_degenerate_batches(ctx):
    import torch
    d = torch.zeros((4, 3 * 44), device="cuda"); o = torch.full((4, 40), 7, dtype=torch.uint8, device="cuda")
    assert ctx.lib.turbo_decode_batch(ctx.h, dptr(d), 40, 0, 0, 0, dptr(o), None) == 0      # n_cb = 0: nothing to do, nothing touched
    assert bool((o == 7).all())
    for mode in (0, 1):                                                                     # all-zero LLRs (max_value = 0 -> NaN -> 0 as on x86)
        ctx.check(ctx.lib.turbo_decode_batch(ctx.h, dptr(d), 40, 4, mode, 2, dptr(o), None))
        torch.cuda.synchronize()
        assert bool((o == 0).all())
    s = torch.full((2, 3 * 44), 10000.0, device="cuda")                                     # everything punctured
    ctx.check(ctx.lib.turbo_decode_batch(ctx.h, dptr(s), 40, 2, 0, 0, dptr(o[:2]), None))
    torch.cuda.synchronize()
    assert bool((o[:2] == 0).all())


def test_utilities_gold_sequence_and_crc_check(ctx, port):
    """lteb200_prs_c_batch (generate_prs_c) and lteb200_crc24a_check_batch (calc_crc) against the oracle."""
    import torch
    import ctypes as C
    from util import dptr, stream_ptr
    for c_init, n in ((0, 100), ((0x1234 << 14) | (7 << 9) | 301, 20000), (0x7FFFFFFF, 31)):
        d = torch.zeros(n, dtype=torch.uint8, device="cuda")
        ctx.check(ctx.lib.prs_c_batch(ctx.h, C.c_uint32(c_init), n, dptr(d), stream_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), port.prs_c(c_init, n))
    rng = np.random.default_rng(2)
    for tbs in (16, 144, 1000, 6120):
        n, stride = 37, tbs + 24 + 5
        blocks = np.zeros((n, stride), np.uint8)
        blocks[:, :tbs] = rng.integers(0, 2, (n, tbs))
        want = np.ones(n, np.int32)
        for b in range(n):
            crc = port.crc(blocks[b, :tbs])
            blocks[b, tbs:tbs + 24] = [(crc >> (23 - i)) & 1 for i in range(24)]
            if b % 3 == 1:
                blocks[b, rng.integers(0, tbs + 24)] ^= 1; want[b] = 0
        d = torch.from_numpy(blocks).cuda(); ok = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        ctx.check(ctx.lib.crc24a_check_batch(ctx.h, dptr(d), stride, tbs, n, dptr(ok), stream_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(ok.cpu().numpy(), want), tbs

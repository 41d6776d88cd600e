"""The CPU half of the max-log-MAP pinning: the decoder's integer definition (oracle/port/turbo_mlm.cc, which the -m gpu tests show to be
bit-identical to both CUDA kernels for all 188 block sizes) against the REFERENCE's turbo_decode (liblte_phy.cc:10620-10845) on code
words from the reference's own encoder -- the same grid and the same two properties as tests/test_turbo_pinning_gpu.py, with 128 blocks
per point, so that the relation kernel == twin (GPU run) and twin <-> reference (every CPU run) are both on record."""
import numpy as np
import pytest

from test_turbo_pinning_gpu import ref_codewords, ref_decode_all


@pytest.mark.parametrize("K", [40, 512, 2048, 3072, 3584, 5248, 6144])
def test_max_log_map_definition_pinned_to_reference(port, refo, K):
    n = 128
    for sigma in (0.0, 0.2, 0.3, 0.4, 0.5):
        cs, xs = ref_codewords(refo, port, K, n, sigma, seed=7000 + K + int(100 * sigma))
        ref_bits = ref_decode_all(refo, port, xs)
        got = np.stack([port.turbo_decode_mlm(x, 8) for x in xs])
        ref_ok = (ref_bits == cs).all(axis=1)
        b_ok = (got == cs).all(axis=1)
        # wherever the reference recovered the block, the max-log-MAP definition produces the same bits ...
        assert np.array_equal(got[ref_ok], ref_bits[ref_ok]), f"K={K} sigma={sigma}"
        # ... and it never does worse
        assert (~b_ok).sum() <= (~ref_ok).sum(), f"K={K} sigma={sigma}: BLER {(~b_ok).mean():.3f} > reference {(~ref_ok).mean():.3f}"
        if sigma <= 0.2:
            assert ref_ok.all()
        if sigma == 0.0:
            assert b_ok.all()

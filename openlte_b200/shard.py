"""Sharding of a long I/Q stream over ranks: contiguous ranges of subframes per rank, no data-path
collective.  The only coupling between neighbouring subframes is the estimator's 2-symbol look-ahead
(liblte_phy_get_dl_subframe_and_ce reads symbols 0-1 of the next subframe, liblte_phy.cc:5946-5957), so each
shard's sample range carries a halo behind its last subframe."""

SF_SAMPS = {0: 1920, 1: 3840, 2: 7680, 3: 15360, 4: 30720}


def halo_samples(fs_idx):
    """two OFDM symbols (first one with the long CP) + the one-sample-early window start"""
    n = 128 << fs_idx
    return 2 * n + (10 << fs_idx) + (9 << fs_idx) + 1


def shard_range(n_sf_total, rank, world):
    """[lo, hi) subframe indices of this rank: contiguous, sizes differ by at most one"""
    base, rem = divmod(n_sf_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_samples(n_sf_total, rank, world, fs_idx, stream_start=0):
    """sample range [s_lo, s_hi) this rank must hold, and the per-subframe start offsets relative to s_lo"""
    lo, hi = shard_range(n_sf_total, rank, world)
    n = SF_SAMPS[fs_idx]
    s_lo = stream_start + lo * n
    s_hi = stream_start + hi * n + halo_samples(fs_idx)
    starts = [i * n for i in range(hi - lo)]
    return s_lo, s_hi, starts


def subframe_numbers(first_subfr, lo, hi):
    return [(first_subfr + i) % 10 for i in range(lo, hi)]

"""One I/Q stream received on N GPUs (BASELINE config 4; the loop being sharded is LTE_fdd_dl_fs_samp_buf.cc:491-515, whose subframes carry no PHY
state from one to the next -- SURVEY 8e).

Rank 0 owns the stream (the "input file").  Per call:
  1. the stream parameters (length, first subframe number, cell) are broadcast,
  2. the samples are distributed -- `broadcast`: the whole int8 stream to every rank (what BASELINE.json's north star names), or `scatter`: each rank
     receives only its contiguous range of subframes plus the estimator's two-symbol halo (openlte_b200/shard.py),
  3. every rank runs the receive chain on its range (no inter-GPU dependency),
  4. transport-block bits and CRC verdicts are gathered on rank 0 in stream order.
The collectives are torch.distributed's (NCCL on GPUs; the CPU test runs the same code on gloo with a stand-in decoder)."""
import numpy as np
import torch
import torch.distributed as dist

from . import shard


def receive_sharded(iq, params, decode_fn, out_shape, device, distribute="broadcast", group=None):
    """iq: rank 0: int8 tensor [n_samps, 2] on `device`; other ranks: None.
    params: rank 0: dict(n_sf, first_subfr, fs_idx) (+ anything decode_fn needs is closed over); other ranks: None.
    decode_fn(iq_local [n, 2] int8, starts (list of sample offsets), subfr_nums (list)) -> (tb uint8 [n_local, *out_shape], ok int32 [n_local, out_shape[0]])
    Returns on rank 0: (tb [n_sf, *out_shape], ok [n_sf, out_shape[0]]); elsewhere (None, None)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    hdr = torch.zeros(4, dtype=torch.int64, device=device)
    if rank == 0:
        hdr[0], hdr[1], hdr[2], hdr[3] = int(iq.shape[0]), int(params["n_sf"]), int(params["first_subfr"]), int(params["fs_idx"])
    dist.broadcast(hdr, 0, group=group)
    n_samps, n_sf, first_subfr, fs_idx = (int(x) for x in hdr.tolist())
    lo, hi = shard.shard_range(n_sf, rank, world)

    def span(r):
        """sample range of rank r, clipped to the stream; (0, 0) for a rank without subframes (more ranks than subframes)"""
        l, h = shard.shard_range(n_sf, r, world)
        if h == l:
            return 0, 0
        a, b, _ = shard.shard_samples(n_sf, r, world, fs_idx)
        return min(a, n_samps), min(b, n_samps)

    s_lo, s_hi = span(rank)
    starts = shard.shard_samples(n_sf, rank, world, fs_idx)[2]
    if distribute == "broadcast":
        buf = iq if rank == 0 else torch.empty((n_samps, 2), dtype=torch.int8, device=device)
        dist.broadcast(buf, 0, group=group)
        mine = buf[s_lo:s_hi]
    else:
        spans = [span(r) for r in range(world)]
        width = max(1, max(b - a for a, b in spans))
        mine_pad = torch.zeros((width, 2), dtype=torch.int8, device=device)
        if rank == 0:
            parts = []
            for a, b in spans:
                p = torch.zeros((width, 2), dtype=torch.int8, device=device)
                p[: b - a] = iq[a:b]
                parts.append(p)
            dist.scatter(mine_pad, parts, src=0, group=group)
        else:
            dist.scatter(mine_pad, None, src=0, group=group)
        mine = mine_pad[: s_hi - s_lo]
    if hi > lo:
        tb, ok = decode_fn(mine, starts, shard.subframe_numbers(first_subfr, lo, hi))
    else:       # nothing to do on this rank; it still takes part in the gather
        tb = torch.zeros((0,) + tuple(out_shape), dtype=torch.uint8, device=device)
        ok = torch.zeros((0, out_shape[0]), dtype=torch.int32, device=device)
    # gather in stream order (shards differ by at most one subframe: pad to the largest)
    sizes = [shard.shard_range(n_sf, r, world) for r in range(world)]
    pad = max(1, max(h - l for l, h in sizes))
    tb_pad = torch.zeros((pad,) + tuple(out_shape), dtype=torch.uint8, device=device)
    ok_pad = torch.zeros((pad, out_shape[0]), dtype=torch.int32, device=device)
    tb_pad[: hi - lo] = tb
    ok_pad[: hi - lo] = ok
    if rank == 0:
        tbs = [torch.empty_like(tb_pad) for _ in range(world)]
        oks = [torch.empty_like(ok_pad) for _ in range(world)]
        dist.gather(tb_pad, tbs, dst=0, group=group)
        dist.gather(ok_pad, oks, dst=0, group=group)
        return (torch.cat([t[: h - l] for t, (l, h) in zip(tbs, sizes)]), torch.cat([o[: h - l] for o, (l, h) in zip(oks, sizes)]))
    dist.gather(tb_pad, None, dst=0, group=group)
    dist.gather(ok_pad, None, dst=0, group=group)
    return None, None

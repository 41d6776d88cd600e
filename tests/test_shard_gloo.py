"""N > 1 host logic on CPU: world_size-2 gloo job; every rank takes its contiguous subframe range (+ halo), "decodes"
it (here: a deterministic function of the samples), and rank 0 gathers the per-subframe results in order."""
import os
import socket
import subprocess
import sys
import textwrap

from openlte_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ranges_cover_everything_once():
    for n in (1, 7, 10, 4000, 10001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    s_lo, s_hi, starts = shard.shard_samples(10, 1, 2, 4)
    assert s_lo == 5 * 30720 and s_hi == 10 * 30720 + shard.halo_samples(4) and starts[1] == 30720
    assert shard.halo_samples(4) >= 2048 + 160 + 2048 + 144       # symbols 14-15 of the last subframe fit
    assert shard.subframe_numbers(1, 8, 12) == [9, 0, 1, 2]


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from openlte_b200 import shard
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_sf, fs_idx = 13, 0
    rng = np.random.default_rng(0)
    stream = rng.integers(-100, 100, (n_sf * 1920 + shard.halo_samples(fs_idx), 2)).astype(np.int8)   # every rank could read the file
    s_lo, s_hi, starts = shard.shard_samples(n_sf, rank, world, fs_idx)
    mine = stream[s_lo:s_hi]
    lo, hi = shard.shard_range(n_sf, rank, world)
    # stand-in for the receive chain: per-subframe checksum over the subframe AND its halo (proves the halo travelled)
    res = torch.tensor([int(mine[s:s + 1920 + shard.halo_samples(fs_idx)].astype(np.int64).sum()) for s in starts], dtype=torch.int64)
    sizes = [shard.shard_range(n_sf, r, world) for r in range(world)]
    pad = max(h - l for l, h in sizes)
    buf = torch.zeros(pad, dtype=torch.int64); buf[:res.numel()] = res
    out = [torch.zeros(pad, dtype=torch.int64) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank == 0:
        got = torch.cat([o[:h - l] for o, (l, h) in zip(out, sizes)]).numpy()
        want = np.array([int(stream[i * 1920:(i + 1) * 1920 + shard.halo_samples(fs_idx)].astype(np.int64).sum()) for i in range(n_sf)])
        assert np.array_equal(got, want), (got, want)
        print("GATHER_OK")
    dist.destroy_process_group()
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "GATHER_OK" in r.stdout


WORKER2 = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from openlte_b200 import shard, sharded
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_sf, fs_idx, first = 11, 0, 7
    halo = shard.halo_samples(fs_idx)
    stream = None
    if rank == 0:
        rng = np.random.default_rng(1)
        stream = torch.from_numpy(rng.integers(-100, 100, (n_sf * 1920 + halo, 2)).astype(np.int8))
    def decode(iq_local, starts, sn):      # stand-in for the receive chain: per-subframe checksum over the subframe AND its halo, tagged with the subframe number
        tb = torch.zeros((len(starts), 2, 4), dtype=torch.uint8); ok = torch.zeros((len(starts), 2), dtype=torch.int32)
        for i, s in enumerate(starts):
            c = int(iq_local[s:s + 1920 + halo].to(torch.int64).sum())
            tb[i, 0] = torch.tensor([(c >> (8 * k)) & 255 for k in range(4)], dtype=torch.uint8)
            tb[i, 1, 0] = sn[i]
            ok[i] = torch.tensor([1, sn[i]])
        return tb, ok
    for mode in ("broadcast", "scatter"):
        tb, ok = sharded.receive_sharded(stream, dict(n_sf=n_sf, first_subfr=first, fs_idx=fs_idx) if rank == 0 else None, decode, (2, 4), "cpu", distribute=mode)
        if rank == 0:
            want_tb, want_ok = decode(stream, [i * 1920 for i in range(n_sf)], shard.subframe_numbers(first, 0, n_sf))
            assert torch.equal(tb, want_tb) and torch.equal(ok, want_ok), mode
    # more ranks than subframes (rank 1 holds nothing, still joins the collectives), and a stream that ends inside the last halo
    short = stream[: 1920 + 100] if rank == 0 else None
    for mode in ("broadcast", "scatter"):
        tb, ok = sharded.receive_sharded(short, dict(n_sf=1, first_subfr=3, fs_idx=fs_idx) if rank == 0 else None, decode, (2, 4), "cpu", distribute=mode)
        if rank == 0:
            want_tb, want_ok = decode(short, [0], [3])
            assert tuple(tb.shape) == (1, 2, 4) and torch.equal(tb, want_tb) and torch.equal(ok, want_ok), mode
        else:
            assert tb is None and ok is None
    if rank == 0:
        print("SHARDED_OK")
    dist.destroy_process_group()
""")


def test_receive_sharded_world_size_2_gloo(tmp_path):
    """openlte_b200/sharded.py (parameter broadcast, stream broadcast / scatter with halo, gather in stream order) equals the single-process result."""
    script = tmp_path / "worker2.py"
    script.write_text(WORKER2 % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout

"""CPU, world_size 2 over gloo: the multi-GPU host logic (channel sharding + all-gather of the
spectrogram columns).  The per-rank compute is stood in for by the CPU oracle here -- the GPU
kernels themselves are covered by the -m gpu tests; this covers the N>1 plumbing."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_everything():
    from friture_b200.sharded import shard_range, shard_sizes
    for C in (1, 7, 8, 256, 1000, 8192):
        for W in (1, 2, 3, 4, 8):
            rs = [shard_range(C, W, r) for r in range(W)]
            assert rs[0][0] == 0 and rs[-1][1] == C
            assert all(rs[i][1] == rs[i + 1][0] for i in range(W - 1))
            assert max(shard_sizes(C, W)) - min(shard_sizes(C, W)) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, n_channels, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from friture_b200.sharded import allgather_channels, shard_range
    from oracle import friture_oracle as fo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)                      # same full input on every rank
        x = (rng.standard_normal((n_channels, 2048 + 3 * 1024)) * 0.1).astype(np.float32)
        lo, hi = shard_range(n_channels, world, rank)
        local = fo.log_spectrogram(fo.stft_power_batch(x[lo:hi], 2048, 1024))
        full = allgather_channels(torch.from_numpy(local), n_channels)
        ref = fo.log_spectrogram(fo.stft_power_batch(x, 2048, 1024))
        ok = bool(np.array_equal(full.numpy(), ref)) and tuple(full.shape) == ref.shape
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_channels", [6, 5])     # even and uneven shards
def test_allgather_columns_world2(n_channels):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_channels, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]

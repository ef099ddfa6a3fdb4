"""Channel sharding across the GPUs of one box (one process per GPU, torch.distributed).

Every channel is an independent stream: the reference processes exactly one channel per call
(friture/spectrogram.py:153-157, friture/octavespectrum.py:97-101), so channels are split into
contiguous blocks, one per rank, with replicated constants and NO exchange while computing.  The
only collective is the optional all-gather of the finished spectrogram columns / band vectors
(NCCL over NVLink on GPUs; gloo on CPU for the host-logic tests).
"""
from __future__ import annotations


def shard_range(n_channels: int, world: int, rank: int):
    """Contiguous channel block [lo, hi) of `rank`; the first n_channels % world ranks get one
    extra channel."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank %d/%d" % (rank, world))
    base, extra = divmod(n_channels, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_sizes(n_channels: int, world: int):
    return [shard_range(n_channels, world, r)[1] - shard_range(n_channels, world, r)[0]
            for r in range(world)]


def allgather_channels(local, n_channels: int, group=None, out=None, async_op=False):
    """All-gather per-rank results along the channel axis (axis 0).

    local: [C_rank, ...] tensor (CUDA for nccl, CPU for gloo); returns [n_channels, ...] on every
    rank in global channel order.  Uneven shards are padded to the largest shard for the
    collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if out is not None:
            out.copy_(local)
            return out
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_channels, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d channels, expected %d" % (rank, local.shape[0], sizes[rank]))
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    if len(set(sizes)) == 1:
        if out is None:
            out = torch.empty((n_channels,) + tail, dtype=local.dtype, device=local.device)
        work = dist.all_gather_into_tensor(out, local, group=group, async_op=async_op)
        return (out, work) if async_op else out
    cmax = max(sizes)
    padded = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
    padded[:sizes[rank]] = local
    buf = torch.empty((world * cmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    if out is None:
        out = torch.empty((n_channels,) + tail, dtype=local.dtype, device=local.device)
    lo = 0
    for r, n in enumerate(sizes):
        out[lo:lo + n] = buf[r * cmax:r * cmax + n]
        lo += n
    return (out, None) if async_op else out


class ShardedSpectrogram:
    """STFT log-power columns of `n_channels` streams sharded over the ranks of the default process
    group.  `process(x_local)` computes this rank's [C_rank, frames, bins] block on its GPU;
    `gather(block)` returns the full [n_channels, frames, bins] array on every rank."""

    def __init__(self, n_channels, n_fft=2048, hop=1024, log=True, handle=None):
        import torch.distributed as dist
        from .audioproc import audioproc
        self.n_channels = n_channels
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.lo, self.hi = shard_range(n_channels, self.world, self.rank)
        self.hop = hop
        self.log = log
        self.proc = audioproc(handle)
        self.proc.set_fftsize(n_fft)

    def process(self, x_local, out=None):
        if x_local.shape[0] != self.hi - self.lo:
            raise ValueError("rank %d expects channels [%d, %d)" % (self.rank, self.lo, self.hi))
        return self.proc.stft(x_local, hop=self.hop, log=self.log, out=out)

    def gather(self, block, out=None):
        return allgather_channels(block, self.n_channels, out=out)

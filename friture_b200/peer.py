"""All-gather over NVLink peer memory: one-hop pushes by the copy engines or by a small copy kernel.

The north-star's one collective is the final all-gather of the spectrogram columns: every GPU must
receive (N-1)/N of ALL columns, about 3.8 GB per GPU and step at 8 x 1024 channels -- link-bound by
construction.  NCCL moves that data with copy KERNELS, whose CTAs share the SMs with the
latency-bound filterbank kernel and take its issue slots (measured at N = 2: +1.2 ms on a 3.2 ms
step for 0.7 ms of link time).  Here every rank allocates its gathered buffer through the library
(``frt_peer_alloc``: a whole ``cudaMalloc`` allocation, exportable over CUDA IPC), the ranks exchange
the 64-byte handles, open each other's buffers, and every rank PUSHES its own block into its peers'
buffers with ``cudaMemcpyAsync`` on one stream per peer: the copy engines drive NVLink/NVSwitch, no
SM is involved, and the transfers overlap the compute for free.  The copy engines of one GPU top out
near 430-510 GB/s (measured at N = 8, where 7/8 of 4.3 GB must leave every GPU per step: 8.66 ms; at
N = 4 they still keep up: 3.17 ms against 3.08 ms of compute), so above 4 GPUs the push is done by
``frt_peer_push`` instead: 64 CTAs that read the block once from local HBM and
store it to all peers with 128-bit stores -- one hop, no ring steps, a few warps' worth of issue
slots (measured at N = 8, 1024 channels x 128 hops per GPU, tools/perf_gather.py: 6.15 ms per step and
611 GB/s received per GPU, against 7.08 ms with NCCL and 8.66 ms with the copy engines).  The rank's own block is written in place by the STFT kernel (``local(i)`` is a view of the
gathered buffer).
"""
from __future__ import annotations

import ctypes
from ctypes import c_size_t, c_void_p

from . import _lib


class _DevBuf:
    """A raw device pointer exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerGather:
    """gathered[i, r] (shape `block_shape`, float32) = block i of rank r, on every rank.

    local(i)            this rank's slot of block i (write the result there)
    push(i, after)      copy block i to every peer once stream `after` has produced it
    join(stream)        make `stream` wait for all of this rank's pushes
    wait_all()          join + device sync + group barrier: every rank's data has landed here
    """

    def __init__(self, handle, n_blocks, block_shape, group=None, engine="auto", n_ctas=64):
        import torch
        import torch.distributed as dist
        self.handle = handle
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_blocks = int(n_blocks)
        self.block_shape = tuple(int(v) for v in block_shape)
        self.block_elems = 1
        for v in self.block_shape:
            self.block_elems *= v
        self.block_bytes = self.block_elems * 4
        total = self.n_blocks * self.world * self.block_bytes
        self.device = torch.device("cuda", handle.device)
        p = c_void_p()
        handle.call("frt_peer_alloc", c_size_t(total), ctypes.byref(p))
        self._ptr = p.value
        self.gathered = torch.as_tensor(_DevBuf(self._ptr, (self.n_blocks, self.world) + self.block_shape),
                                        device=self.device)
        # "ce": one cudaMemcpyAsync per peer (copy engines, no SM at all; 430-510 GB/s per GPU measured:
        # enough up to 4 GPUs, where the step stays at the compute time); "kernel": frt_peer_push, a
        # small copy kernel that reads the block once and stores it to every peer (611 GB/s at 8 GPUs;
        # at 4 GPUs its CTAs cost the filterbank more than the copy engines' lower rate: 4.76 vs 3.17 ms)
        self.engine = ("ce" if self.world <= 4 else "kernel") if engine == "auto" else engine
        self.n_ctas = int(n_ctas)
        self._peers = [None] * self.world
        self._streams = [None] * self.world
        self._push_stream = None
        if self.world > 1:
            mine = ctypes.create_string_buffer(64)
            handle.call("frt_peer_export", c_void_p(self._ptr), mine)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine.raw, group=group)
            for r in range(self.world):
                if r == self.rank:
                    continue
                q = c_void_p()
                handle.call("frt_peer_import", ctypes.create_string_buffer(handles[r], 64), ctypes.byref(q))
                self._peers[r] = q.value
                self._streams[r] = torch.cuda.Stream(self.device)
            self._push_stream = torch.cuda.Stream(self.device, priority=-1)
            order = [(self.rank + k) % self.world for k in range(1, self.world)]   # stagger the targets
            self._order = order
            self._peer_table = (c_void_p * len(order))()
            dist.barrier(group=group)

    def local(self, i):
        return self.gathered[i, self.rank]

    def push(self, i, after):
        """Queue the copies of this rank's block i into every peer's buffer behind stream `after`."""
        off = (i * self.world + self.rank) * self.block_bytes
        src = c_void_p(self._ptr + off)
        if self.world > 1 and self.engine == "kernel":
            for n, r in enumerate(self._order):
                self._peer_table[n] = self._peers[r] + off
            st = self._push_stream
            st.wait_stream(after)
            self.handle.call("frt_peer_push", src, self._peer_table, len(self._order), c_size_t(self.block_bytes),
                             self.n_ctas, c_void_p(st.cuda_stream))
            return
        for k in range(1, self.world):
            r = (self.rank + k) % self.world         # stagger the targets over the ranks
            st = self._streams[r]
            st.wait_stream(after)
            self.handle.call("frt_peer_copy", c_void_p(self._peers[r] + off), src, c_size_t(self.block_bytes),
                             c_void_p(st.cuda_stream))

    def join(self, stream):
        for st in self._streams + [self._push_stream]:
            if st is not None:
                stream.wait_stream(st)

    def wait_all(self):
        import torch
        import torch.distributed as dist
        self.join(torch.cuda.current_stream(self.device))
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def close(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self._peers):
            if p is not None:
                self.handle.call("frt_peer_close", c_void_p(p))
                self._peers[r] = None
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)      # nobody frees a buffer a peer still has open
        if self._ptr:
            self.gathered = None
            self.handle.call("frt_peer_free", c_void_p(self._ptr))
            self._ptr = None

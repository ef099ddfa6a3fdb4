"""Multi-GPU step time of the combined workload by gather transport (run under torchrun).

    python -m torch.distributed.run --nproc-per-node N tools/perf_gather.py [--chunks 8]

Prints one line per transport: max-over-ranks ms per step (CUDA events, barrier on both sides).
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from friture_b200.analyzer import ChannelAnalyzer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--chunks", type=int, nargs="+", default=[8])
    ap.add_argument("--ctas", type=int, nargs="+", default=[16, 32, 64, 128])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    C, F = args.channels, args.frames
    x = torch.randn((C, (F + 1) * 1024), device=dev)
    bands = torch.empty((C, F + 1, 30), device=dev)
    an = ChannelAnalyzer(C, 2048, 1024, 3, 10, device=dev.index)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        if getattr(an, "peer_gather", None) is not None:
            an.peer_gather.join(torch.cuda.current_stream())
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for nch in args.chunks:
        fc = F // nch
        configs = [("nccl", "auto", 0), ("peer", "ce", 0)] + [("peer", "kernel", n) for n in args.ctas]
        for transport, engine, ctas in configs:
            gathered = spec_chunks = None
            if transport == "nccl":
                spec_chunks = torch.empty((nch, C, fc, 1025), device=dev)
                gathered = torch.empty((nch, world * C, fc, 1025), device=dev)

            def fn():
                an.process_sharded(x, gathered, spec_chunks, bands, nch, transport=transport, engine=engine)
            if transport == "peer":
                fn()
                an.peer_gather.n_ctas = ctas or 32
            ms = timed(fn)
            if rank == 0:
                rx = (world - 1) * C * F * 1025 * 4 / (ms * 1e-3) / 1e9
                print("chunks %2d %-5s %-6s ctas %3d: %.3f ms/step  %.0f GB/s received per GPU  %.3g units/s"
                      % (nch, transport, engine, ctas, ms, rx, world * C * F / (ms * 1e-3)), flush=True)
            del gathered, spec_chunks
    if getattr(an, "peer_gather", None) is not None:
        an.peer_gather.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

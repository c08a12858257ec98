#!/usr/bin/env python
"""Does an NCCL all-gather on a side stream overlap a compute kernel on the main stream — eagerly and inside a CUDA graph?
torchrun --nproc-per-node N tools/overlap_nccl_probe.py"""
import json
import os

import torch
import torch.distributed as dist


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n = 13015680
    table = torch.randn(n, device=dev).half()
    lo = rank * (n // world)
    x = torch.randn(1 << 24, device=dev)
    side = torch.cuda.Stream()

    def compute():                      # ~0.15-0.2 ms of SM-bound work on the main stream
        y = x
        for _ in range(6):
            y = torch.sin(y)
        return y

    def gather():
        dist.all_gather_into_tensor(table, table[lo:lo + n // world])

    def overlapped():
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            gather()
            ev = torch.cuda.Event()
            ev.record(side)
        compute()
        main.wait_event(ev)

    def sequential():
        gather()
        compute()

    res = {"world": world, "compute_ms": timeit(compute), "gather_ms": timeit(gather), "eager_sequential_ms": timeit(sequential),
           "eager_overlapped_ms": timeit(overlapped)}
    for name, fn in (("graph_sequential_ms", sequential), ("graph_overlapped_ms", overlapped)):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        res[name] = timeit(g.replay)
        g = None
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

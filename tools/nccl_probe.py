#!/usr/bin/env python
"""Device time of the collectives the gradient exchange uses, at the sizes it uses them (torchrun --nproc-per-node N tools/nccl_probe.py)."""
import json
import os

import torch
import torch.distributed as dist


def timeit(fn, iters=30):
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
    g32, g16 = torch.randn(n, device=dev), torch.randn(n, device=dev).half()
    sh16, sh32 = torch.empty(n // world, device=dev, dtype=torch.half), torch.empty(n // world, device=dev)
    small, flag = torch.randn(18432, device=dev), torch.zeros(1, device=dev)
    res = {"world": world, "elements": n}
    res["all_reduce_fp32_52MB_ms"] = timeit(lambda: dist.all_reduce(g32))
    res["all_reduce_fp16_26MB_ms"] = timeit(lambda: dist.all_reduce(g16))
    res["reduce_scatter_fp32_ms"] = timeit(lambda: dist.reduce_scatter_tensor(sh32, g32))
    res["reduce_scatter_fp16_ms"] = timeit(lambda: dist.reduce_scatter_tensor(sh16, g16))
    lo = rank * (n // world)
    res["all_gather_fp16_ms"] = timeit(lambda: dist.all_gather_into_tensor(g16, g16[lo:lo + n // world]))
    res["all_reduce_74KB_ms"] = timeit(lambda: dist.all_reduce(small))
    res["all_reduce_4B_ms"] = timeit(lambda: dist.all_reduce(flag, op=dist.ReduceOp.MAX))
    res["cast_fp32_to_fp16_ms"] = timeit(lambda: g16.copy_(g32))
    res["isfinite_any_shard_ms"] = timeit(lambda: (~torch.isfinite(sh16)).any())
    res["mul_fp32_ms"] = timeit(lambda: g32.mul_(0.5))
    # the same collectives replayed from a CUDA graph (what the bench does)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            dist.reduce_scatter_tensor(sh16, g16)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        dist.reduce_scatter_tensor(sh16, g16)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        dist.all_reduce(small)
    res["graph_rs16_flag_small_ms"] = timeit(gr.replay)
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    gr = None
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Where does the sharded gradient exchange spend its time?  Replays its pieces from CUDA graphs on a table-sized dummy model.
torchrun --nproc-per-node N tools/exchange_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import parallel  # noqa: E402
from enerf_b200.gridencoder.grid import _half_table  # noqa: E402
from enerf_b200.optim import FusedAdam  # noqa: E402


def graph_time(fn, iters=40):
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
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}
    for mode in ("sharded", "allreduce"):
        table = torch.nn.Parameter(torch.randn(6507840, 2, device=dev) * 0.01)
        w1, w2 = torch.nn.Parameter(torch.randn(7168, device=dev)), torch.nn.Parameter(torch.randn(11264, device=dev))
        model = torch.nn.ParameterList([table, w1, w2])
        opt = FusedAdam([{"params": [table]}, {"params": [w1]}, {"params": [w2]}], lr=1e-3, betas=(0.9, 0.99), eps=1e-15)
        ex = (parallel.ShardedExchange if mode == "sharded" else parallel.AllReduceExchange)(model, opt)
        with torch.autocast("cuda", dtype=torch.float16):
            _half_table(table)
        grads = [torch.randn_like(p) for p in (table, w1, w2)]
        opt.grad_scale = torch.full((1,), 1024.0, device=dev)
        opt.found_inf = torch.zeros(1, device=dev)

        def set_grads():
            for p, g in zip((table, w1, w2), grads):
                p.grad = g

        set_grads()
        ex.before_step()
        opt.step()
        ex.after_step()
        res[mode + "_before_step_ms"] = graph_time(lambda: (set_grads(), ex.before_step()))
        res[mode + "_optimizer_ms"] = graph_time(lambda: opt.step())

        def whole():
            ex.begin_step()
            _half_table(table)
            set_grads()
            ex.before_step()
            opt.step()
            ex.after_step()

        res[mode + "_begin+before+opt_ms"] = graph_time(whole)
        if mode == "sharded":
            res["sharded_gather_alone_ms"] = graph_time(lambda: (setattr(ex, "_dirty", True), ex.begin_step(), ex._wait_table()))
            from enerf_b200 import _lib
            wire, flag = torch.empty(table.numel(), dtype=torch.half, device=dev), torch.zeros(1, device=dev)
            res["grad_to_half_ms"] = graph_time(lambda: _lib.call("enerf_grad_to_half", _lib.ptr(grads[0]), _lib.ptr(wire), table.numel(), 8000.0, _lib.ptr(flag),
                                                                    _lib.stream()))
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-kernel device timing on realistic inputs (the bench.py workload's marched samples).
Usage: python tools/bench_kernels.py [--rays 4096] [--iters 10] [--only grid,mlp,march,...]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import _lib, synthetic  # noqa: E402
from enerf_b200 import raymarching as rm  # noqa: E402
from enerf_b200.backends import ffmlp_backend as FB, gridencoder_backend as GB  # noqa: E402
from enerf_b200.gridencoder import GridEncoder  # noqa: E402


def timeit(fn, iters, flush=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--bound", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    dev = torch.device("cuda", 0)
    bound = args.bound
    cascade = 1 + int(np.ceil(np.log2(bound)))
    bits = torch.from_numpy(synthetic.packbits_np(synthetic.ball_density_grid(bound, cascade))).to(dev)
    o, d = synthetic.random_rays(args.rays, bound, seed=100)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, float(bound), bits, cascade, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
    S = xyzs.shape[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2
    res = {"samples": S, "rays": args.rays}
    peak = 6545.0

    def want(k):
        return not only or k in only

    if want("march"):
        M = S
        bx, bd, bdl = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        br = torch.empty(args.rays, 3, dtype=torch.int32, device=dev)

        def f():
            counter.zero_()
            _lib.call("enerf_march_rays_train", o.data_ptr(), d.data_ptr(), bits.data_ptr(), float(bound), 0.0, 1024, args.rays, cascade, 128, M,
                      nears.data_ptr(), fars.data_ptr(), bx.data_ptr(), bd.data_ptr(), bdl.data_ptr(), br.data_ptr(), counter.data_ptr(), 1, _lib.stream())
        res["march_ms"] = timeit(f, args.iters, flush)

    if want("grid"):
        enc = GridEncoder(desired_resolution=2048 * bound).to(dev)
        x01 = ((xyzs + bound) / (2 * bound)).contiguous()
        for dt in (torch.float16, torch.float32):
            emb = enc.embeddings.detach().to(dt).contiguous()
            out = torch.empty(S, 32, device=dev, dtype=dt)
            dummy = torch.empty(1, device=dev, dtype=dt)
            Sx = float(np.log2(enc.per_level_scale))
            med, mn = timeit(lambda: GB.grid_encode_forward(x01, emb, enc.offsets, out, S, 3, 2, 16, Sx, 16, False, dummy, 0, 1), args.iters, flush)
            nbytes = (12 + 16 * 8 * 2 * emb.element_size() + 32 * emb.element_size()) * S
            res[f"grid_fwd_{str(dt)[6:]}"] = {"ms": med, "min_ms": mn, "GBps": nbytes / med / 1e6, "frac": nbytes / med / 1e6 / peak}
            grad = torch.randn(S, 32, device=dev).to(dt)
            for gdt in ((torch.float32, torch.float16) if dt == torch.float16 else (torch.float32,)):
                for mode in (1, 0):
                    _lib.call("enerf_grid_set_backward_mode", mode)
                    gg = torch.zeros(emb.shape, device=dev, dtype=gdt)
                    med, mn = timeit(lambda: GB.grid_encode_backward(grad, x01, emb, enc.offsets, gg, S, 3, 2, 16, Sx, 16, False, dummy, dummy, 0, 1),
                                     args.iters, flush)
                    nbytes = (12 + 32 * emb.element_size() + 2 * 16 * 8 * 2 * emb.element_size()) * S
                    res[f"grid_bwd_{str(dt)[6:]}_acc{str(gdt)[6:]}_{'walk' if mode else 'percorner'}"] = {
                        "ms": med, "min_ms": mn, "GBps": nbytes / med / 1e6, "frac": nbytes / med / 1e6 / peak}
                _lib.call("enerf_grid_set_backward_mode", 1)
        # random (incoherent) points for comparison
        xr = torch.rand(S, 3, device=dev)
        emb = enc.embeddings.detach().half().contiguous()
        out = torch.empty(S, 32, device=dev, dtype=torch.half)
        dummy = torch.empty(1, device=dev, dtype=torch.half)
        med, mn = timeit(lambda: GB.grid_encode_forward(xr, emb, enc.offsets, out, S, 3, 2, 16, Sx, 16, False, dummy, 0, 1), args.iters, flush)
        res["grid_fwd_float16_randompts"] = {"ms": med, "GBps": 588 * S / med / 1e6}

    if want("mlp"):
        for nl, name in ((2, "sigma"), (3, "colour")):
            nw = 64 * (32 + 64 * (nl - 1) + 16)
            w = ((torch.rand(nw, device=dev) * 2 - 1) * (3 / 64) ** 0.5).half()
            x = (torch.randn(S, 32, device=dev) * 0.5).half()
            g = (torch.randn(S, 16, device=dev) * 0.1).half()
            out = torch.empty(S, 16, device=dev, dtype=torch.half)
            fb = torch.empty(nl, S, 64, device=dev, dtype=torch.half)
            bb = torch.empty(nl, S, 64, device=dev, dtype=torch.half)
            gi = torch.empty(S, 32, device=dev, dtype=torch.half)
            gw = torch.empty(nw, device=dev, dtype=torch.float32)
            flops_f = 2.0 * nw * S
            for path in (0, 1):
                _lib.call("enerf_ffmlp_set_path", path)
                tag = f"mlp_{name}_{'tc' if path == 0 else 'mma'}"
                med, mn = timeit(lambda: FB.ffmlp_forward(x, w, S, 32, 16, 64, nl, 0, 6, fb, out), args.iters, flush)
                res[tag + "_fwd"] = {"ms": med, "min_ms": mn, "TFLOPs": flops_f / med / 1e9, "frac": flops_f / med / 1e9 / 1387.0}
                med, mn = timeit(lambda: FB.ffmlp_inference(x, w, S, 32, 16, 64, nl, 0, 6, None, out), args.iters, flush)
                res[tag + "_inf"] = {"ms": med, "min_ms": mn, "TFLOPs": flops_f / med / 1e9, "frac": flops_f / med / 1e9 / 1387.0}
                med, mn = timeit(lambda: FB.ffmlp_backward(g, x, w, fb, S, 32, 16, 64, nl, 0, 6, True, bb, gi, gw), args.iters, flush)
                res[tag + "_bwd"] = {"ms": med, "min_ms": mn, "TFLOPs": 2 * flops_f / med / 1e9, "frac": 2 * flops_f / med / 1e9 / 1387.0}
                if path == 0:   # what training runs: no backward_buffer (activation gradients stay on the SM; TMA operand loads)
                    med, mn = timeit(lambda: FB.ffmlp_backward(g, x, w, fb, S, 32, 16, 64, nl, 0, 6, True, None, gi, gw), args.iters, flush)
                    res[tag + "_bwd_nobuf"] = {"ms": med, "min_ms": mn, "TFLOPs": 2 * flops_f / med / 1e9, "frac": 2 * flops_f / med / 1e9 / 1387.0}
                    # no forward_buffer either: hidden activations recomputed per tile (what the fused field trains with)
                    med, mn = timeit(lambda: FB.ffmlp_backward(g, x, w, None, S, 32, 16, 64, nl, 0, 6, True, None, gi, gw), args.iters, flush)
                    res[tag + "_bwd_recompute"] = {"ms": med, "min_ms": mn, "TFLOPs": 2 * flops_f / med / 1e9, "frac": 2 * flops_f / med / 1e9 / 1387.0}
            _lib.call("enerf_ffmlp_set_path", 0)

    if want("composite"):
        sig = torch.rand(S, device=dev) * 10
        rgb = torch.rand(S, 1, device=dev)
        N = args.rays
        ws, dp, im = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 1, device=dev)
        med, mn = timeit(lambda: _lib.call("enerf_composite_rays_train_forward", sig.data_ptr(), rgb.data_ptr(), deltas.data_ptr(), rays.data_ptr(), S, N, 1,
                                           ws.data_ptr(), dp.data_ptr(), im.data_ptr(), _lib.stream()), args.iters, flush)
        res["composite_fwd"] = {"ms": med, "GBps": 16.0 * S / med / 1e6}

    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

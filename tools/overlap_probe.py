#!/usr/bin/env python
"""Does the hash-grid kernel overlap with the MLP kernels when they run on two streams?

The scatter is bound by L2 reductions and the gather by L1/LSU sectors, the tcgen05 MLP kernels by the shared-memory /
tensor pipes (backward) and HBM (sigma forward): different bottlenecks, so chunk k of one can run next to chunk k+1 of the
other.  This probe times, on the bench workload's marched samples,

  bwd: colour-net bwd -> sigma-net bwd -> scatter        sequential vs `tools/pipelined_backward.py`
  fwd: gather -> sigma-net fwd -> colour-net fwd         sequential vs the same pipeline in the other direction

for a few (chunks, MLP CTA cap, scatter CTA size) settings and checks that the results agree.
Usage: python tools/overlap_probe.py [--rays 4096] [--iters 10] [--out gpurun_out/overlap_probe.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import _lib, synthetic  # noqa: E402
from tools import pipelined_backward as pipelined  # noqa: E402
from enerf_b200 import raymarching as rm  # noqa: E402
from enerf_b200._lib import ptr, stream  # noqa: E402
from enerf_b200.backends import gridencoder_backend as GB  # noqa: E402
from enerf_b200.gridencoder import GridEncoder  # noqa: E402


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(float(np.median(ts)), 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--bound", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    bound = args.bound
    cascade = 1 + int(np.ceil(np.log2(bound)))
    bits = torch.from_numpy(synthetic.packbits_np(synthetic.ball_density_grid(bound, cascade))).to(dev)
    o, d = synthetic.random_rays(args.rays, bound, seed=100)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, _, _ = rm.march_rays_train(o, d, float(bound), bits, cascade, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
    S = xyzs.shape[0] // 128 * 128
    x = ((xyzs[:S] + bound) / (2 * bound)).contiguous()
    dirs = dirs[:S].contiguous()

    torch.manual_seed(0)
    enc = GridEncoder(desired_resolution=2048 * bound).to(dev)
    table = (torch.rand_like(enc.embeddings) - 0.5).half().contiguous()
    offsets = enc.offsets
    geometry = (S, 3, 2, 16, np.log2(enc.per_level_scale), enc.base_resolution)
    nl_s, nl_c, n_ch = 2, 3, 1
    ws = ((torch.rand(64 * (32 + 64 * (nl_s - 1) + 16), device=dev) * 2 - 1) * (3 / 64) ** 0.5).half()
    wc = ((torch.rand(64 * (32 + 64 * (nl_c - 1) + 16), device=dev) * 2 - 1) * (3 / 64) ** 0.5).half()
    feat = torch.empty(S, 32, dtype=torch.half, device=dev)
    sigma = torch.empty(S, device=dev)
    cin = torch.empty(S, 32, dtype=torch.half, device=dev)
    rgb = torch.empty(S, n_ch, device=dev)
    dummy = table.new_empty(1)
    side = pipelined._side_stream(dev)

    def gather(lo, hi):
        GB.grid_encode_forward(x[lo:hi], table, offsets, feat[lo:hi], hi - lo, 3, 2, 16, geometry[4], geometry[5], False, dummy, 0, 1)

    def mlp_fwd(lo, hi):
        _lib.call("enerf_field_sigma_forward", ptr(feat[lo:hi]), ptr(ws), ptr(dirs[lo:hi]), hi - lo, nl_s, None, ptr(sigma[lo:hi]), ptr(cin[lo:hi]), stream())
        _lib.call("enerf_field_color_forward", ptr(cin[lo:hi]), ptr(wc), hi - lo, nl_c, n_ch, None, ptr(rgb[lo:hi]), None, stream())

    def fwd_seq():
        gather(0, S)
        mlp_fwd(0, S)

    def fwd_pipe(chunks, cap):
        step = -(-(S // 128) // chunks) * 128
        main = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        _lib.call("enerf_ffmlp_set_max_ctas", cap)
        try:
            for lo in range(0, S, step):
                hi = min(S, lo + step)
                with torch.cuda.stream(side):
                    gather(lo, hi)
                    ev = torch.cuda.Event()
                    ev.record(side)
                main.wait_event(ev)
                mlp_fwd(lo, hi)
        finally:
            _lib.call("enerf_ffmlp_set_max_ctas", 0)

    res = {"samples": S, "rays": args.rays, "fwd": [], "bwd": []}
    fwd_seq()
    torch.cuda.synchronize()
    ref_sigma, ref_rgb, ref_feat = sigma.clone(), rgb.clone(), feat.clone()
    res["fwd_sequential_ms"] = timeit(fwd_seq, args.iters)
    res["gather_ms"] = timeit(lambda: gather(0, S), args.iters)
    res["mlp_fwd_ms"] = timeit(lambda: mlp_fwd(0, S), args.iters)
    for chunks, cap in [(2, 148), (4, 148), (4, 132), (4, 120), (4, 104), (8, 120), (8, 104), (16, 120)]:
        try:
            sigma.zero_(), rgb.zero_(), feat.zero_()
            fwd_pipe(chunks, cap)
            torch.cuda.synchronize()
            same = bool(torch.equal(sigma, ref_sigma) and torch.equal(rgb, ref_rgb) and torch.equal(feat, ref_feat))
            res["fwd"].append({"chunks": chunks, "mlp_ctas": cap, "ms": timeit(lambda: fwd_pipe(chunks, cap), args.iters), "bit_identical": same})
        except Exception as e:  # noqa: BLE001
            res["fwd"].append({"chunks": chunks, "mlp_ctas": cap, "error": f"{type(e).__name__}: {e}"})
        print(json.dumps(res["fwd"][-1]), flush=True)

    # ---------------- backward
    sigma.copy_(ref_sigma), rgb.copy_(ref_rgb), feat.copy_(ref_feat)
    g_sigma = torch.randn(S, device=dev) * 1e-3
    g_rgb = torch.randn(S, n_ch, device=dev) * 1e-2
    dcin = torch.empty(S, 32, dtype=torch.half, device=dev)
    dfeat = torch.empty(S, 32, dtype=torch.half, device=dev)
    gw_c = torch.empty(wc.numel(), device=dev)
    gw_s = torch.empty(ws.numel(), device=dev)
    d_table = torch.empty(table.shape, dtype=torch.float32, device=dev)

    def mlp_bwd():
        _lib.call("enerf_field_color_backward", ptr(g_rgb), ptr(rgb), n_ch, ptr(cin), ptr(wc), None, S, nl_c, ptr(dcin), ptr(gw_c), None, stream())
        _lib.call("enerf_field_sigma_backward", ptr(g_sigma), ptr(sigma), ptr(dcin), ptr(feat), ptr(ws), None, S, nl_s, ptr(dfeat), ptr(gw_s), stream())

    def scatter():
        d_table.zero_()
        GB.grid_encode_backward(dfeat, x, table, offsets, d_table, S, 3, 2, 16, geometry[4], geometry[5], False, dummy, dummy, 0, 1)

    def bwd_seq():
        mlp_bwd()
        scatter()

    bwd_seq()
    torch.cuda.synchronize()
    ref_table, ref_gws, ref_gwc = d_table.clone(), gw_s.clone(), gw_c.clone()
    res["bwd_sequential_ms"] = timeit(bwd_seq, args.iters)
    res["mlp_bwd_ms"] = timeit(mlp_bwd, args.iters)
    res["scatter_ms"] = timeit(scatter, args.iters)

    def bwd_pipe(chunks, cap, block):
        return pipelined.pipelined_backward(g_sigma, g_rgb, sigma, rgb, cin, feat, x, table, offsets, geometry, 0, ws, wc, nl_s, nl_c, n_ch,
                                        torch.float32, chunks, cap, block)

    for chunks, cap, block in [(1, 148, 256), (2, 148, 256), (4, 148, 256), (4, 148, 128), (4, 132, 256), (4, 120, 256), (4, 104, 256), (4, 88, 256),
                               (4, 120, 128), (8, 120, 256), (8, 104, 256), (8, 132, 128), (16, 120, 256)]:
        try:
            t_, s_, c_ = bwd_pipe(chunks, cap, block)
            torch.cuda.synchronize()
            err = {"table": float((t_ - ref_table).abs().max() / ref_table.abs().max()), "gw_sigma": float((s_ - ref_gws).abs().max() / ref_gws.abs().max()),
                   "gw_color": float((c_ - ref_gwc).abs().max() / ref_gwc.abs().max())}
            res["bwd"].append({"chunks": chunks, "mlp_ctas": cap, "scatter_block": block, "ms": timeit(lambda: bwd_pipe(chunks, cap, block), args.iters),
                               "max_err_rel_to_max": err})
        except Exception as e:  # noqa: BLE001
            res["bwd"].append({"chunks": chunks, "mlp_ctas": cap, "scatter_block": block, "error": f"{type(e).__name__}: {e}"})
        print(json.dumps(res["bwd"][-1]), flush=True)

    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""The fused inference field (csrc/field_infer.cu) against the chain it replaces: (a) one call on the marcher-ordered samples of the bench
workload (4096 rays, ~3.29 M samples): grid_encode_forward + field_sigma_forward + field_color_forward vs enerf_field_infer;
(b) the 800x800 frame of BASELINE configs[3] through both.  (While the kernel was being shaped this script also walked its
instantiations — MLP slots x gather teams x level pairs in flight; those numbers are in profiles/r2_50_field_infer_probe.json.)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from enerf_b200 import raymarching, synthetic  # noqa: E402


def time_ms(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    frames = "--no-frame" not in sys.argv
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    D = bench.Dist()
    torch.manual_seed(0)
    model = bench.make_ff_model(dev, bench.BOUND)
    model.eval()
    out = {"per_call": {}, "frame": {}}

    # (a) the samples of one training march (what the bench's kernels table is quoted on)
    o_np, d_np = synthetic.random_rays(4096, bench.BOUND, seed=4242)
    o, d = torch.from_numpy(o_np).to(dev), torch.from_numpy(d_np).to(dev)
    nears, fars = raymarching.near_far_from_aabb(o, d, model.aabb_train, model.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(o, d, model.bound, model.density_bitfield, model.cascade, model.grid_size, nears, fars,
                                                            counter, 1 << 20, False, 128, True, 0, 1024)
    n = xyzs.shape[0]                                  # the emitted samples, padded to a multiple of 128 with zero rows
    xyzs, dirs = xyzs.contiguous(), dirs.contiguous()
    out["per_call"]["samples"] = n
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        model.fuse_infer = False
        ref = model(xyzs, dirs)
        out["per_call"]["unfused_chain_ms"] = time_ms(lambda: model(xyzs, dirs))
        model.fuse_infer = True
        got = model(xyzs, dirs)
        out["per_call"]["fused_ms"] = time_ms(lambda: model(xyzs, dirs))
        out["per_call"]["bit_identical"] = bool(torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]))

    # (b) the full frame
    if frames:
        for key, fuse in (("unfused", False), ("fused", True)):
            model.fuse_infer = fuse
            r = bench.render_bench(model, dev, D, frames=2)
            out["frame"][key] = {"frame_ms": r["frame_ms"], "msamples_per_s": r["msamples_per_s"], "samples": r["samples_shaded"]}
    print(json.dumps({"field_infer_probe": out}))


if __name__ == "__main__":
    main()

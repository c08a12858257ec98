#!/usr/bin/env python
"""Inference gather vs sample order.  One marching round of an 800 x 800 frame (26 steps per ray) is encoded in (a) the order march_rays
emits (ray-major: a warp's 32 samples are consecutive steps of ONE ray), (b) step-major (a warp = the same step of 32 consecutive
pixels of an image row), (c) tile-major (a warp = the same step of an 8 x 4 pixel tile)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import synthetic  # noqa: E402
from enerf_b200 import raymarching as rm  # noqa: E402
from enerf_b200.gridencoder import GridEncoder  # noqa: E402


def timeit(fn, iters=7):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda", 0)
    bound, W, n_step = 3, 800, 26
    grid = synthetic.ball_density_grid(bound, 3)
    bits = torch.from_numpy(synthetic.packbits_np(grid)).to(dev)
    pose = synthetic.look_at_poses(1, 0.6 * bound, seed=5)[0]
    o, d = synthetic.pinhole_rays(pose, W, W)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    N = o.shape[0]
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
    rays_t = nears.clone()
    res = {"rays": N, "n_step": n_step, "rounds": []}
    enc = GridEncoder(desired_resolution=2048 * bound).to(dev)
    with torch.no_grad():
        enc.embeddings.uniform_(-0.5, 0.5)
    for rnd in range(3):
        xyzs, dirs, deltas = rm.march_rays(N, n_step, rays_alive, rays_t, o, d, float(bound), bits, 3, 128, nears, fars, -1, False, 0, 1024)
        x = xyzs[:N * n_step].view(N, n_step, 3)
        valid = float((deltas[:N * n_step, 0] > 0).float().mean())
        orders = {"ray_major": x.reshape(-1, 3)}
        orders["step_major"] = x.transpose(0, 1).contiguous().view(-1, 3)
        t = x.view(W // 4, 4, W // 8, 8, n_step, 3).permute(0, 2, 4, 1, 3, 5).contiguous().view(-1, 3)      # (tile row, tile col, step, 4 x 8 pixels)
        orders["tile_major"] = t
        row = {"round": rnd, "valid_fraction": valid}
        for name, pts in orders.items():
            pts = pts.contiguous()
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                row[name + "_ms"] = timeit(lambda: enc(pts, bound=bound))
        res["rounds"].append(row)
        rays_t = rays_t + deltas[:N * n_step, 1].view(N, n_step).sum(dim=1)       # every ray advances by what it marched
    print(json.dumps(res))


if __name__ == "__main__":
    main()

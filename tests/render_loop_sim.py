#!/usr/bin/env python
"""How many rows does the inference loop of NeRFRenderer.run_cuda (enerf_b200/nerf/renderer.py, reference nerf/renderer.py:364-391)
hand to the field, and how many of them belong to rays that are still alive?

CPU replay of the loop's bookkeeping on the bench's 800 x 800 frame (BASELINE configs[3]: analytic-ball occupancy, camera at 0.6 * bound):
the per-ray sample counts come from the oracle's marcher on every 8th pixel of every 8th row (10 000 rays, scaled to 640 000), a ray
is alive until a round leaves it fewer samples than the round's n_step.  The host sizes a round by the last alive count it has read
(every `sync_every` rounds); the true count is what the device-side counter holds.  Cited by DESIGN.md 3.10
(`enerf_field_infer_alive`): with sync_every = 4 the host-sized rounds contain 5.8 % more rows than live ones.

    python tests/render_loop_sim.py            # test infrastructure: loads oracle/, nothing of it is on the product path
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import synthetic  # noqa: E402
from oracle import oracle  # noqa: E402

BOUND, RES, STRIDE, BUDGET = 3, 800, 8, 1 << 24


def per_ray_samples():
    pose = synthetic.look_at_poses(1, 0.6 * BOUND, seed=7)[0]
    ys, xs = np.meshgrid(np.arange(0, RES, STRIDE), np.arange(0, RES, STRIDE), indexing="ij")
    pix = (ys * RES + xs).reshape(-1)
    o, d = synthetic.pinhole_rays(pose, RES, RES, 50.0, pix)
    cascade = 3
    bits = synthetic.packbits_np(synthetic.ball_density_grid(BOUND, cascade))
    aabb = np.array([-BOUND] * 3 + [BOUND] * 3, np.float32)
    nears, fars = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    _, _, _, rays, _ = oracle.march_rays_train(o, d, BOUND, bits, cascade, 128, nears, fars, perturb=False)
    k = np.zeros(len(pix), np.int64)
    k[rays[:, 0]] = rays[:, 2]
    return k


def replay(k, sync_every):
    n_rays = RES * RES
    scale = n_rays / len(k)
    left, alive = k.copy(), np.ones(len(k), bool)
    n_bound, true_count, step, i, since = n_rays, n_rays, 0, 0, 0
    rows_host, rows_live = 0, 0
    while step < 1024:
        if i > 0:
            true_count = int(alive.sum() * scale)
            since += 1
            if since >= sync_every:
                n_bound, since = min(n_bound, true_count), 0
        if n_bound <= 0:
            break
        n_step = max(min(n_rays // n_bound, 8), 1)
        n_step = max(n_step, min(BUDGET // n_bound, 1024 - step))
        rows_host += n_bound * n_step
        rows_live += true_count * n_step
        emitted = np.minimum(left, n_step)
        died = alive & (emitted < n_step)
        left = left - np.where(alive, emitted, 0)
        alive &= ~died
        step += n_step
        i += 1
    return {"sync_every": sync_every, "rounds": i, "rows_sized_by_the_host": int(rows_host), "rows_of_live_rays": int(rows_live),
            "excess": rows_host / rows_live - 1.0}


def main():
    k = per_ray_samples()
    out = {"rays_sampled": int(len(k)), "mean_samples_per_ray": float(k.mean()), "max_samples_per_ray": int(k.max()),
           "policies": [replay(k, s) for s in (1, 2, 4, 8)]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE configs[2]-shaped step (mocapDesk2: event_only + accumulate_evs, C_thres = -1 normalised loss, 8192 event pairs/batch,
bound 2) through every piece of this repository: device event-pair sampler (N3) -> event rays + near/far (N2) -> two renders of
the cuda_ray / ff model (the hot path) -> fused event loss (N1) -> backward -> GradScaler + FusedAdam (N4).

  python tools/bench_event_step.py [--pairs 8192] [--steps 50] [--warmup 5]
Prints one JSON line: event pairs/s and rendered rays/s (2 rays per pair), ms/step, samples per step.  Synthetic event frame
(events grouped by pixel, per-event poses = a look-at pose jittered by 0.2 deg / 1 mm, polarities +-1).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import events, synthetic  # noqa: E402
from enerf_b200.graphs import GraphedStep  # noqa: E402
from enerf_b200.nerf.network_ff import NeRFNetwork  # noqa: E402
from enerf_b200.optim import FusedAdam  # noqa: E402

BOUND = 2
H_EV, W_EV = 260, 346


def synthetic_event_frame(n_pixels=60000, seed=0):
    rng = np.random.default_rng(seed)
    counts = rng.integers(2, 12, n_pixels)
    pix = rng.choice(H_EV * W_EV, n_pixels, replace=False)
    xs = np.repeat(pix % W_EV, counts).astype(np.float32)
    ys = np.repeat(pix // W_EV, counts).astype(np.float32)
    E = int(counts.sum())
    ts = rng.random(E).astype(np.float32)
    pol = rng.choice([-1.0, 1.0], E).astype(np.float32)
    ev = np.stack([xs, ys, ts, pol], 1)
    cum = np.cumsum(counts)
    start = np.repeat(cum - counts, counts)
    num_succ = (np.repeat(cum, counts) - np.arange(E) - 1).astype(np.int64)
    del start
    base = synthetic.look_at_poses(1, 0.6 * BOUND, seed=3)[0]
    # per-event pose: small rotation about a random axis (sigma 0.2 deg) and translation (sigma 1 mm)
    ang = np.radians(0.2) * rng.normal(size=(E, 3)).astype(np.float32)
    K = np.zeros((E, 3, 3), np.float32)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ang[:, 2], ang[:, 1], ang[:, 2], -ang[:, 0], -ang[:, 1], ang[:, 0]
    R = (np.eye(3, dtype=np.float32)[None] + K) @ (base[:3, :3] * np.array([1, -1, -1], np.float32))       # OpenCV-style camera axes
    t = base[:3, 3][None] + 1e-3 * rng.normal(size=(E, 3)).astype(np.float32)
    poses = np.concatenate([R, t[:, :, None]], -1).astype(np.float32)
    return ev, num_succ, cum - 1, poses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--graph", default="on", choices=["on", "off"])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = NeRFNetwork(encoding="hashgrid", bound=BOUND, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1,
                        out_dim_color=1).to(dev).train()
    grid = synthetic.ball_density_grid(BOUND, model.cascade)
    model.density_grid.copy_(torch.from_numpy(grid))
    model.density_bitfield.copy_(torch.from_numpy(synthetic.packbits_np(grid)))
    ev, num_succ, no_succ, poses = synthetic_event_frame()
    sampler = events.EventPairSampler(ev, num_succ, no_succ, acc_max_num_evs=8, poses_evs=poses, device=dev)
    focal = H_EV / (2 * np.tan(np.radians(50.0) / 2))
    intr = (focal, focal, W_EV / 2, H_EV / 2)
    optimizer = FusedAdam(model.get_params(5e-3), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    kw = dict(num_steps=512, upsample_steps=0, max_ray_batch=5096, dt_gamma=0, out_dim_color=1)
    bg = torch.rand(1, 1, 1, device=dev)

    def step():
        batch = sampler.sample(a.pairs, intr)
        with torch.autocast("cuda", dtype=torch.float16):
            out1 = model.render(batch["rays_evs_o1"], batch["rays_evs_d1"], staged=False, bg_color=bg, perturb=True, **kw)
            out2 = model.render(batch["rays_evs_o2"], batch["rays_evs_d2"], staged=False, bg_color=bg, perturb=True, **kw)
        loss, _ = events.event_loss(out1["image"].float(), out2["image"].float(), batch["pols"], use_luma=False, linlog=True, C_thres=-1,
                                    event_only=True)
        optimizer.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
        return loss

    step()                                              # sizes the sample buffers
    totals = model.step_counter[:2, 0].tolist()
    model.mean_count = int(1.03 * max(totals))
    run = step
    if a.graph == "on":
        g = GraphedStep(lambda: step(), [], warmup=3)
        run = lambda: g()                               # noqa: E731
    for _ in range(max(3, a.warmup)):
        loss = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"workload": "BASELINE configs[2]-shaped: event_only + accumulate_evs, normalised loss (C_thres=-1), bound 2, ff + cuda_ray, fp16 autocast; "
                                  "sampler (N3) -> event rays (N2) -> 2 renders -> event loss (N1) -> backward -> GradScaler + FusedAdam (N4)",
                      "event_pairs_per_step": a.pairs, "ms_per_step": ms, "event_pairs_per_s": a.pairs / (ms * 1e-3), "rendered_rays_per_s": 2 * a.pairs / (ms * 1e-3),
                      "samples_per_render": [int(t) for t in totals], "launch": "cuda-graph replay" if a.graph == "on" else "eager",
                      "loss": float(loss), "loss_finite": bool(torch.isfinite(loss))}))


if __name__ == "__main__":
    main()

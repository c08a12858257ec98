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
H_EV, W_EV = synthetic.EVENT_H, synthetic.EVENT_W


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
    ev, num_succ, no_succ, poses = synthetic.event_frame(bound=BOUND)
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

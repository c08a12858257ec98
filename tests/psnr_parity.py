#!/usr/bin/env python
"""PSNR parity on the tiny synthetic scene (BASELINE configs[0] / SURVEY.md §8d C1).

Two trainings from identical initial parameters on identical ray batches:
  reference : CPU port of the reference's pure-PyTorch path (nerf/network.py + NeRFRenderer.run), oracle/cpu_reference.py
  ours      : enerf_b200.nerf.network.NeRFNetwork on the GPU (same topology; hash grid, SH, near/far and the run() integrator
              are this repo's CUDA kernels), fp32, no autocast
then both render the 8 training poses; reports PSNR vs the analytic target for each, their difference (the north star asks
for <= 0.1 dB) and the PSNR between the two renderings.

  python tests/psnr_parity.py [--steps 200] [--num-steps 128] [--rays 256] [--res 64]
Test infrastructure (it imports oracle/): used by tests/test_gpu_renderer.py::test_psnr_parity_tiny_scene.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import synthetic  # noqa: E402


def scene(n_poses=8, res=64, bound=1, ball=0.35):
    """rays of `n_poses` cameras at 0.6*bound and the analytic target: a Lambert-shaded coloured ball on white."""
    poses = synthetic.look_at_poses(n_poses, 0.6 * bound, seed=1)
    os_, ds_ = [], []
    for p in poses:
        o, d = synthetic.pinhole_rays(p, res, res, 50.0)
        os_.append(o)
        ds_.append(d)
    o, d = np.concatenate(os_), np.concatenate(ds_)
    b = (o * d).sum(-1)
    c = (o * o).sum(-1) - (ball * bound) ** 2
    disc = b * b - c
    hit = disc > 0
    t = -b - np.sqrt(np.where(hit, disc, 0))
    hit &= t > 0
    n = (o + t[:, None] * d) / (ball * bound)
    light = np.array([0.5, 0.7, 0.5]) / np.linalg.norm([0.5, 0.7, 0.5])
    shade = np.clip((n * light).sum(-1), 0.1, 1.0)
    base = 0.5 + 0.5 * n
    rgb = np.where(hit[:, None], base * shade[:, None], 1.0).astype(np.float32)
    return o, d, rgb


def psnr(a, b):
    return float(-10 * np.log10(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2) + 1e-20))


def run(steps=200, num_steps=128, rays=256, res=64, bound=1, lr=5e-3, seed=0, verbose=False):
    from enerf_b200.nerf.network import NeRFNetwork
    from oracle import cpu_reference
    dev = torch.device("cuda", 0)
    o, d, rgb = scene(res=res, bound=bound)
    torch.manual_seed(seed)
    ref = cpu_reference.NeRFNetworkCPU(bound=bound, out_dim_color=3)
    ours = NeRFNetwork(encoding="hashgrid", bound=bound, cuda_ray=False, out_dim_color=3).to(dev)
    with torch.no_grad():                                      # identical initial parameters
        ours.encoder.embeddings.copy_(ref.encoder.embeddings)
        for a, b in zip(list(ours.sigma_net) + list(ours.color_net), list(ref.sigma_net) + list(ref.color_net)):
            a.weight.copy_(b.weight)
    assert ours.encoder.offsets.cpu().tolist() == ref.encoder.offsets.tolist()
    opt_r = torch.optim.Adam(ref.parameters(), lr=lr, betas=(0.9, 0.99), eps=1e-15)
    opt_o = torch.optim.Adam(ours.parameters(), lr=lr, betas=(0.9, 0.99), eps=1e-15)
    rng = np.random.default_rng(seed)
    to, td, tt = torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(rgb)
    go, gd, gt = to.to(dev), td.to(dev), tt.to(dev)
    ours.train()
    for it in range(steps):
        idx = torch.from_numpy(rng.integers(0, len(o), size=rays))
        out_r = ref.render(to[idx], td[idx], num_steps=num_steps, perturb=False)
        loss_r = F.mse_loss(out_r["image"], tt[idx])
        opt_r.zero_grad(set_to_none=True)
        loss_r.backward()
        opt_r.step()
        gi = idx.to(dev)
        out_o = ours.render(go[gi].unsqueeze(0), gd[gi].unsqueeze(0), staged=False, bg_color=1, perturb=False, num_steps=num_steps,
                            upsample_steps=0)
        loss_o = F.mse_loss(out_o["image"].reshape(-1, 3), gt[gi])
        opt_o.zero_grad(set_to_none=True)
        loss_o.backward()
        opt_o.step()
        if verbose and (it % 25 == 0 or it == steps - 1):
            print(f"step {it:4d}  loss ref {float(loss_r):.6f}  ours {float(loss_o):.6f}", flush=True)
    ours.eval()
    img_r, img_o = [], []
    with torch.no_grad():
        for s in range(0, len(o), 4096):
            img_r.append(ref.render(to[s:s + 4096], td[s:s + 4096], num_steps=num_steps, perturb=False)["image"].numpy())
            img_o.append(ours.render(go[s:s + 4096].unsqueeze(0), gd[s:s + 4096].unsqueeze(0), staged=False, bg_color=1, perturb=False,
                                     num_steps=num_steps, upsample_steps=0)["image"].reshape(-1, 3).cpu().numpy())
    img_r, img_o = np.concatenate(img_r), np.concatenate(img_o)
    return {"psnr_reference_db": psnr(img_r, rgb), "psnr_ours_db": psnr(img_o, rgb), "psnr_between_db": psnr(img_r, img_o),
            "abs_diff_db": abs(psnr(img_r, rgb) - psnr(img_o, rgb)), "final_loss_reference": float(loss_r), "final_loss_ours": float(loss_o),
            "steps": steps, "num_steps": num_steps, "rays_per_batch": rays, "poses": 8, "res": res,
            "config": "tiny synthetic scene (8 poses, 64x64 RGB), nerf/network.py topology, run() (no cuda_ray, no ff), fp32, perturb off"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--num-steps", type=int, default=128)
    ap.add_argument("--rays", type=int, default=256)
    ap.add_argument("--res", type=int, default=64)
    a = ap.parse_args()
    print(json.dumps(run(a.steps, a.num_steps, a.rays, a.res, verbose=True)))

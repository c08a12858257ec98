#!/usr/bin/env python
"""Training-level parity of the benchmarked hot path against the reference's OWN CUDA kernels (north star: "rendered PSNR within
0.1 dB of the reference").

Two trainings of the tiny synthetic scene (8 poses, 64x64 RGB, analytic shaded ball) from identical initial parameters, on identical
ray batches, fp16 autocast + GradScaler, occupancy grid refreshed every 16 steps (nerf/utils.py:945-947), `max_steps` 1024:
  reference : oracle/ref_chain.RefStack — nerf/network_ff.py + the training branch of NeRFRenderer.run_cuda chained on the reference's
              unmodified extensions built for sm_100a (oracle/_ref): wmma/CUTLASS MLP with fp16 accumulation, fp16-atomic hash-grid
              backward, thread-per-ray marcher and compositor; torch.optim.Adam
  ours      : enerf_b200.nerf.network_ff.NeRFNetwork — warp-per-ray marcher, gather, tcgen05 fused field (recomputing backward),
              walking scatter, FusedAdam
Both parameter sets are then rendered with the same inference renderer (this repo's, eval mode) on the 8 training poses: PSNR vs the
analytic target, their difference, and the PSNR between the two renderings.  `gradient_check()` compares the composed gradient of
one training step (same bitfield, same samples) between the two stacks.

  python tests/hotpath_parity.py [--steps 200] [--rays 1024] [--out profiles/x.json]
Test infrastructure (imports oracle/): used by tests/test_gpu_hotpath_parity.py."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.psnr_parity import psnr, scene  # noqa: E402


def _models(bound, dev, seed=0):
    from enerf_b200.nerf.network_ff import NeRFNetwork
    from oracle import ref_chain
    torch.manual_seed(seed)
    ours = NeRFNetwork(encoding="hashgrid", bound=bound, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1,
                       out_dim_color=3).to(dev).train()
    theirs = ref_chain.RefStack(bound=bound).to(dev).train()
    with torch.no_grad():
        theirs.encoder.embeddings.copy_(ours.encoder.embeddings)
        theirs.w_sigma.copy_(ours.sigma_net.weights)
        theirs.w_color.copy_(ours.color_net.weights)
    return ours, theirs


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def gradient_check(n_rays=1024, bound=1, seed=0):
    """one training step on identical samples (the occupancy bitfield of `ours` is given to both; the two marchers are bit-identical):
    image / depth and the gradient of every parameter, tcgen05 stack vs reference kernels"""
    dev = torch.device("cuda", 0)
    ours, theirs = _models(bound, dev, seed)
    with torch.no_grad():
        ours.encoder.embeddings.uniform_(-0.5, 0.5)                # non-degenerate densities
        theirs.encoder.embeddings.copy_(ours.encoder.embeddings)
    torch.manual_seed(1)
    with torch.autocast("cuda", dtype=torch.float16):
        ours.update_extra_state()
    theirs.density_grid.copy_(ours.density_grid)
    theirs.density_bitfield.copy_(ours.density_bitfield)
    o, d, rgb = scene(res=64, bound=bound)
    idx = np.random.default_rng(seed).integers(0, len(o), size=n_rays)
    go, gd, gt = (torch.from_numpy(a[idx]).to(dev) for a in (o, d, rgb))
    scale = 1024.0                                                 # GradScaler's role: fp16 gradients out of the subnormal range
    with torch.autocast("cuda", dtype=torch.float16):
        out_o = ours.render(go[None], gd[None], staged=False, bg_color=1, perturb=True, force_all_rays=True, out_dim_color=3)
    (F.mse_loss(out_o["image"].reshape(-1, 3).float(), gt) * scale).backward()
    out_t = theirs.render_train(go, gd, bg_color=1, perturb=True, force_all_rays=True)
    (F.mse_loss(out_t["image"].float(), gt) * scale).backward()
    samples = (int(ours.step_counter[0, 0]), int(theirs.step_counter[0, 0]))
    res = {"samples_ours": samples[0], "samples_reference": samples[1],
           "image_max_abs_diff": float((out_o["image"].reshape(-1, 3).float() - out_t["image"].float()).abs().max()),
           "depth_max_abs_diff": float((out_o["depth"].reshape(-1) - out_t["depth"]).abs().max()),
           "grad_embeddings_rel_l2": _rel(ours.encoder.embeddings.grad, theirs.encoder.embeddings.grad),
           "grad_sigma_net_rel_l2": _rel(ours.sigma_net.weights.grad, theirs.w_sigma.grad),
           "grad_color_net_rel_l2": _rel(ours.color_net.weights.grad, theirs.w_color.grad),
           "grad_embeddings_norm": float(theirs.encoder.embeddings.grad.norm())}
    return res


def run(steps=600, n_rays=1024, bound=1, lr=5e-3, seed=0, verbose=False, eval_last=5, eval_every=10, grad_accumulation="fp32"):
    """grad_accumulation: 'fp32' = this repo's default; 'fp16' = embedding gradients accumulated with fp16 atomics like the reference
    (gridencoder.cu:296-302), which shows how much of a PSNR difference is the reference's lossy accumulation and not the kernels"""
    from enerf_b200.gridencoder import grid as grid_mod
    grid_mod.set_grad_accumulation(grad_accumulation)
    try:
        return _run(steps, n_rays, bound, lr, seed, verbose, eval_last, eval_every, grad_accumulation)
    finally:
        grid_mod.set_grad_accumulation("fp32")


def _run(steps, n_rays, bound, lr, seed, verbose, eval_last, eval_every, grad_accumulation):
    from enerf_b200.optim import FusedAdam
    dev = torch.device("cuda", 0)
    ours, theirs = _models(bound, dev, seed)
    o, d, rgb = scene(res=64, bound=bound)
    go, gd, gt = (torch.from_numpy(a).to(dev) for a in (o, d, rgb))
    opt_o = FusedAdam(ours.get_params(lr), betas=(0.9, 0.99), eps=1e-15)
    opt_t = torch.optim.Adam(theirs.parameters(), lr=lr, betas=(0.9, 0.99), eps=1e-15)
    sc_o, sc_t = torch.amp.GradScaler("cuda"), torch.amp.GradScaler("cuda")
    # the reference's schedule: lr * 0.1 ** (step / iters) (main_nerf.py:212)
    sch_o = torch.optim.lr_scheduler.LambdaLR(opt_o, lambda it: 0.1 ** min(it / steps, 1))
    sch_t = torch.optim.lr_scheduler.LambdaLR(opt_t, lambda it: 0.1 ** min(it / steps, 1))
    rng = np.random.default_rng(seed)
    log, evals = [], []

    def evaluate():
        """PSNR of both parameter sets on the 8 training poses through the SAME inference renderer (this repo's, eval mode)"""
        was = ours.training
        ours.eval()
        keep = {k: v.detach().clone() for k, v in (("emb", ours.encoder.embeddings), ("ws", ours.sigma_net.weights), ("wc", ours.color_net.weights),
                                                 ("grid", ours.density_grid), ("bits", ours.density_bitfield))}
        other = {"emb": theirs.encoder.embeddings.detach(), "ws": theirs.w_sigma.detach(), "wc": theirs.w_color.detach(), "grid": theirs.density_grid,
                 "bits": theirs.density_bitfield}
        images = {}
        for name, st in (("reference", other), ("ours", keep)):          # `ours` last: its own state is back in place afterwards
            with torch.no_grad():
                ours.encoder.embeddings.copy_(st["emb"])
                ours.sigma_net.weights.copy_(st["ws"])
                ours.color_net.weights.copy_(st["wc"])
                ours.density_grid.copy_(st["grid"])
                ours.density_bitfield.copy_(st["bits"])
                parts = []
                for s0 in range(0, len(o), 8192):
                    with torch.autocast("cuda", dtype=torch.float16):
                        parts.append(ours.render(go[s0:s0 + 8192][None], gd[s0:s0 + 8192][None], staged=False, bg_color=1, perturb=False,
                                                 out_dim_color=3)["image"].reshape(-1, 3).float().cpu().numpy())
                images[name] = np.concatenate(parts)
        ours.train(was)
        return psnr(images["ours"], rgb), psnr(images["reference"], rgb), psnr(images["ours"], images["reference"])

    for it in range(steps):
        if it % 16 == 0:                                           # nerf/utils.py:945-947; same jitter seed for both refreshes
            for m in (ours, theirs):
                torch.manual_seed(1000 + it)
                with torch.autocast("cuda", dtype=torch.float16):
                    m.update_extra_state()
        idx = torch.from_numpy(rng.integers(0, len(o), size=n_rays)).to(dev)
        with torch.autocast("cuda", dtype=torch.float16):
            out_o = ours.render(go[idx][None], gd[idx][None], staged=False, bg_color=1, perturb=True, out_dim_color=3)
        loss_o = F.mse_loss(out_o["image"].reshape(-1, 3).float(), gt[idx])
        opt_o.zero_grad(set_to_none=True)
        sc_o.scale(loss_o).backward()
        sc_o.step(opt_o)
        sc_o.update()
        sch_o.step()
        out_t = theirs.render_train(go[idx], gd[idx], bg_color=1, perturb=True)
        loss_t = F.mse_loss(out_t["image"].float(), gt[idx])
        opt_t.zero_grad(set_to_none=True)
        sc_t.scale(loss_t).backward()
        sc_t.step(opt_t)
        sc_t.update()
        sch_t.step()
        if it >= steps - eval_last * eval_every and (steps - 1 - it) % eval_every == 0:
            evals.append((it,) + evaluate())
        if it % 25 == 0 or it == steps - 1:
            log.append((it, float(loss_o), float(loss_t)))
            if verbose:
                print(f"step {it:4d}  loss ours {float(loss_o):.6f}  reference kernels {float(loss_t):.6f}  samples {int(ours.step_counter[(ours.local_step - 1) % 16, 0])} / "
                      f"{int(theirs.step_counter[(theirs.local_step - 1) % 16, 0])}", flush=True)
    # both trainings are chaotic (different rounding -> diverging trajectories): the PSNR of a single checkpoint wobbles by more than the
    # 0.1 dB the north star asks for, so the last `eval_last` checkpoints (every `eval_every` steps) are averaged
    p_o, p_t = float(np.mean([e[1] for e in evals])), float(np.mean([e[2] for e in evals]))
    return {"psnr_ours_db": p_o, "psnr_reference_kernels_db": p_t, "abs_diff_db": abs(p_o - p_t), "psnr_between_db": float(np.mean([e[3] for e in evals])),
            "checkpoints": [{"step": e[0], "psnr_ours_db": e[1], "psnr_reference_kernels_db": e[2], "psnr_between_db": e[3]} for e in evals],
            "psnr_std_over_checkpoints_db": [float(np.std([e[1] for e in evals])), float(np.std([e[2] for e in evals]))],
            "grad_accumulation_ours": grad_accumulation, "final_loss_ours": log[-1][1], "final_loss_reference_kernels": log[-1][2], "loss_log": log, "steps": steps, "rays_per_batch": n_rays,
            "config": "tiny synthetic scene (8 poses, 64x64 RGB), ff + cuda_ray, fp16 autocast + GradScaler, max_steps 1024, perturb on, "
                      "update_extra_state every 16 steps, lr 5e-3 * 0.1^(step/steps) (main_nerf.py:212); reference side = the reference's own CUDA build (oracle/_ref) chained as network_ff.py + run_cuda"}


def train_single(kind, steps=1500, n_rays=1024, bound=1, lr=5e-3, seed=0, eval_last=5, eval_every=10):
    """one training of ONE stack ('ours' or 'reference'), same protocol as run(); returns the PSNR (mean of the last checkpoints) on the
    8 training poses.  Two calls with the same arguments differ only by what is not reproducible in the stack itself (the order of the
    marcher's atomic reservations and of the gradient atomics): `run_to_run_noise` uses that as the floor of any PSNR comparison."""
    from enerf_b200.optim import FusedAdam
    dev = torch.device("cuda", 0)
    ours, theirs = _models(bound, dev, seed)
    o, d, rgb = scene(res=64, bound=bound)
    go, gd, gt = (torch.from_numpy(a).to(dev) for a in (o, d, rgb))
    model = ours if kind == "ours" else theirs
    opt = FusedAdam(ours.get_params(lr), betas=(0.9, 0.99), eps=1e-15) if kind == "ours" else torch.optim.Adam(theirs.parameters(), lr=lr, betas=(0.9, 0.99),
                                                                                                            eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** min(it / steps, 1))
    rng = np.random.default_rng(seed)
    vals = []
    for it in range(steps):
        if it % 16 == 0:
            torch.manual_seed(1000 + it)
            with torch.autocast("cuda", dtype=torch.float16):
                model.update_extra_state()
        idx = torch.from_numpy(rng.integers(0, len(o), size=n_rays)).to(dev)
        if kind == "ours":
            with torch.autocast("cuda", dtype=torch.float16):
                img = ours.render(go[idx][None], gd[idx][None], staged=False, bg_color=1, perturb=True, out_dim_color=3)["image"].reshape(-1, 3)
        else:
            img = theirs.render_train(go[idx], gd[idx], bg_color=1, perturb=True)["image"]
        loss = F.mse_loss(img.float(), gt[idx])
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        sched.step()
        if it >= steps - eval_last * eval_every and (steps - 1 - it) % eval_every == 0:
            if kind == "reference":                                  # render the reference's parameters with this repo's inference renderer
                with torch.no_grad():
                    ours.encoder.embeddings.copy_(theirs.encoder.embeddings)
                    ours.sigma_net.weights.copy_(theirs.w_sigma)
                    ours.color_net.weights.copy_(theirs.w_color)
                    ours.density_grid.copy_(theirs.density_grid)
                    ours.density_bitfield.copy_(theirs.density_bitfield)
            ours.eval()
            parts = []
            with torch.no_grad():
                for s0 in range(0, len(o), 8192):
                    with torch.autocast("cuda", dtype=torch.float16):
                        parts.append(ours.render(go[s0:s0 + 8192][None], gd[s0:s0 + 8192][None], staged=False, bg_color=1, perturb=False,
                                                 out_dim_color=3)["image"].reshape(-1, 3).float().cpu().numpy())
            ours.train()
            vals.append(psnr(np.concatenate(parts), rgb))
    return float(np.mean(vals))


def run_to_run_noise(kind, pairs=4, steps=1500, n_rays=1024):
    """PSNR of `pairs` x 2 trainings of the same stack from the same seed: |a - b| is pure run-to-run noise"""
    a = np.array([[train_single(kind, steps, n_rays, seed=s) for _ in range(2)] for s in range(pairs)])
    diff = a[:, 0] - a[:, 1]
    return {"stack": kind, "pairs": pairs, "steps": steps, "psnr_db": a.tolist(), "abs_diff_db": np.abs(diff).tolist(),
            "rms_diff_db": float(np.sqrt((diff ** 2).mean())), "max_abs_diff_db": float(np.abs(diff).max())}


def render_parity(steps=400, n_rays=1024, bound=1, seed=0):
    """"rendered PSNR within 0.1 dB of the reference", the deterministic part: ONE trained parameter set (ours, `steps` steps) rendered
    on the 8 training poses by (a) this repo's inference path (device-counted marching loop, tcgen05 field, warp compositor) and (b) the
    reference's own inference kernels chained as renderer.py:344-400 (march_rays / ffmlp_inference / composite_rays / compact_rays of
    oracle/_ref).  Same parameters, same occupancy bitfield, perturb off."""
    from enerf_b200.optim import FusedAdam
    dev = torch.device("cuda", 0)
    ours, theirs = _models(bound, dev, seed)
    o, d, rgb = scene(res=64, bound=bound)
    go, gd, gt = (torch.from_numpy(a).to(dev) for a in (o, d, rgb))
    opt = FusedAdam(ours.get_params(5e-3), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    rng = np.random.default_rng(seed)
    for it in range(steps):
        if it % 16 == 0:
            with torch.autocast("cuda", dtype=torch.float16):
                ours.update_extra_state()
        idx = torch.from_numpy(rng.integers(0, len(o), size=n_rays)).to(dev)
        with torch.autocast("cuda", dtype=torch.float16):
            img = ours.render(go[idx][None], gd[idx][None], staged=False, bg_color=1, perturb=True, out_dim_color=3)["image"].reshape(-1, 3)
        loss = F.mse_loss(img.float(), gt[idx])
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
    with torch.no_grad():
        theirs.encoder.embeddings.copy_(ours.encoder.embeddings)
        theirs.w_sigma.copy_(ours.sigma_net.weights)
        theirs.w_color.copy_(ours.color_net.weights)
        theirs.density_grid.copy_(ours.density_grid)
        theirs.density_bitfield.copy_(ours.density_bitfield)
    ours.eval()
    theirs.eval()
    img_o, img_t, dep_o, dep_t = [], [], [], []
    with torch.no_grad():
        for s0 in range(0, len(o), 8192):
            with torch.autocast("cuda", dtype=torch.float16):
                r = ours.render(go[s0:s0 + 8192][None], gd[s0:s0 + 8192][None], staged=False, bg_color=1, perturb=False, out_dim_color=3)
            img_o.append(r["image"].reshape(-1, 3).float().cpu().numpy())
            dep_o.append(r["depth"].reshape(-1).float().cpu().numpy())
            t = theirs.render_infer(go[s0:s0 + 8192], gd[s0:s0 + 8192], bg_color=1, perturb=False)
            img_t.append(t["image"].float().cpu().numpy())
            dep_t.append(t["depth"].float().cpu().numpy())
    img_o, img_t, dep_o, dep_t = (np.concatenate(a) for a in (img_o, img_t, dep_o, dep_t))
    return {"train_steps": steps, "psnr_ours_db": psnr(img_o, rgb), "psnr_reference_kernels_db": psnr(img_t, rgb),
            "abs_diff_db": abs(psnr(img_o, rgb) - psnr(img_t, rgb)), "psnr_between_db": psnr(img_o, img_t),
            "image_max_abs_diff": float(np.abs(img_o - img_t).max()), "depth_max_abs_diff": float(np.abs(dep_o - dep_t).max()),
            "depth_mean_abs_diff": float(np.abs(dep_o - dep_t).mean()), "pixels": int(img_o.shape[0])}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--out", default="")
    ap.add_argument("--seeds", type=int, default=1, help="independent repetitions (different initial parameters and ray batches)")
    ap.add_argument("--noise", type=int, default=0, help="only measure run-to-run noise: this many same-seed pairs of each stack (negative: render parity of that many trained checkpoints)")
    a = ap.parse_args()
    if a.noise < 0:
        res = {"render_parity": [render_parity(a.steps, a.rays, seed=s_) for s_ in range(-a.noise)]}
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
        sys.exit(0)
    if a.noise > 0:
        res = {"run_to_run_noise": [run_to_run_noise("ours", a.noise, a.steps, a.rays), run_to_run_noise("reference", a.noise, a.steps, a.rays)]}
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
        sys.exit(0)
    res = {"gradient_check": gradient_check(), "training": run(a.steps, a.rays, verbose=True)}
    if a.seeds > 1:
        # both trainings are chaotic and (atomics) not even reproducible run to run: the comparison is between MEANS over repetitions
        runs = [res["training"]] + [run(a.steps, a.rays, seed=s, verbose=False) for s in range(1, a.seeds)]
        po, pt = np.array([r["psnr_ours_db"] for r in runs]), np.array([r["psnr_reference_kernels_db"] for r in runs])
        res["repetitions"] = {"seeds": a.seeds, "steps": a.steps, "psnr_ours_db": po.tolist(), "psnr_reference_kernels_db": pt.tolist(),
                              "mean_ours_db": float(po.mean()), "mean_reference_kernels_db": float(pt.mean()),
                              "mean_diff_db": float((po - pt).mean()), "stderr_of_mean_diff_db": float((po - pt).std(ddof=1) / np.sqrt(a.seeds)),
                              "std_ours_db": float(po.std(ddof=1)), "std_reference_kernels_db": float(pt.std(ddof=1))}
        print("repetitions", res["repetitions"])
        res["training_fp16_grad_accumulation"] = run(a.steps, a.rays, verbose=False, grad_accumulation="fp16")
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)

"""Kernel times of the reference's OWN CUDA extensions (oracle/_ref, built for sm_100a by oracle/build_ref.py from the unmodified
sources) next to this repo's kernels, on the bench workload's marched samples, same B200, same process — "the GPU bar".
BASELINE INFRASTRUCTURE (it loads oracle/_ref): used by bench.py's `gpu_bar` object (outside every timed region) and by
tests/bench_vs_reference_build.py; never on the product path.  The reference kernels launch on the legacy default stream, which
is torch's default stream, so torch events bracket them correctly."""
import json

import numpy as np
import torch

from enerf_b200 import synthetic
from enerf_b200 import raymarching as rm
from enerf_b200.backends import ffmlp_backend as FB, gridencoder_backend as GB, raymarching_backend as RB
from enerf_b200.gridencoder import GridEncoder
from . import ref


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



def measure(n_rays=4096, iters=5, bound=3, out_path="", verbose=False):
    """-> dict(samples, rays, unit, rows=[{kernel, ours_ms, reference_build_ms, speedup, note}], totals)"""
    dev = torch.device("cuda", torch.cuda.current_device())
    cascade = 1 + int(np.ceil(np.log2(bound)))
    bits = torch.from_numpy(synthetic.packbits_np(synthetic.ball_density_grid(bound, cascade))).to(dev)
    o, d = synthetic.random_rays(n_rays, bound, seed=100)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, float(bound), bits, cascade, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
    S = xyzs.shape[0] // 128 * 128
    N = n_rays
    rows = []
    res = {"samples": S, "rays": N, "unit": "ms (median of %d, CUDA events)" % iters, "rows": rows}

    def row(name, ours, theirs, note=""):
        r = {"kernel": name}
        for key, fn in (("ours_ms", ours), ("reference_build_ms", theirs)):
            if fn is None:
                continue
            try:
                r[key] = timeit(fn, iters)
            except Exception as e:  # noqa: BLE001
                r[key + "_error"] = f"{type(e).__name__}: {e}"[:300]
                torch.cuda.synchronize()
        if "ours_ms" in r and "reference_build_ms" in r:
            r["speedup"] = round(r["reference_build_ms"] / r["ours_ms"], 2)
        if note:
            r["note"] = note
        rows.append(r)
        if verbose:
            print(json.dumps(r), flush=True)
        if out_path:
            with open(out_path, "w") as f:
                json.dump(res, f, indent=1)

    # ---------------- marcher (K5)
    R = ref.load("_raymarching")
    M = xyzs.shape[0]
    bx, bd, bdl = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
    br = torch.empty(N, 3, dtype=torch.int32, device=dev)

    def march_ours():
        counter.zero_()
        # as the product calls it (raymarching.march_rays_train): the box around the occupied cells, then the bounded march
        RB.march_rays_train(o, d, bits, float(bound), 0.0, 1024, N, cascade, 128, M, nears, fars, bx, bd, bdl, br, counter, 1,
                            RB.occupancy_bounds(bits, cascade, 128))

    def march_ref():
        counter.zero_()
        R.march_rays_train(o, d, bits, float(bound), 0.0, 1024, N, cascade, 128, M, nears, fars, bx, bd, bdl, br, counter, 1)

    row("march_rays_train", march_ours, march_ref if R else None, "ours: enerf_occupancy_bounds + enerf_march_rays_train_bounded (same samples, bit for bit)")

    # ---------------- compositing (K6/K7), 3 channels as in the reference
    sig = torch.rand(M, device=dev) * 20
    rgb3 = torch.rand(M, 3, device=dev)
    ws, dp, im = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
    g1, g3 = torch.randn(N, device=dev), torch.randn(N, 3, device=dev)
    gs, gr = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
    row("composite_rays_train_forward", lambda: RB.composite_rays_train_forward(sig, rgb3, deltas, rays, M, N, ws, dp, im),
        (lambda: R.composite_rays_train_forward(sig, rgb3, deltas, rays, M, N, ws, dp, im)) if R else None)
    row("composite_rays_train_backward", lambda: RB.composite_rays_train_backward(g1, g3, sig, rgb3, deltas, rays, ws, im, M, N, gs, gr),
        (lambda: R.composite_rays_train_backward(g1, g3, sig, rgb3, deltas, rays, ws, im, M, N, gs, gr)) if R else None)
    del sig, rgb3, gs, gr, bx, bd, bdl

    # ---------------- hash grid (K11/K12), fp16 table as under autocast
    R = ref.load("_gridencoder")
    enc = GridEncoder(desired_resolution=2048 * bound).to(dev)
    table = (torch.rand_like(enc.embeddings) - 0.5).half().contiguous()
    offsets = enc.offsets
    x = ((xyzs[:S] + bound) / (2 * bound)).contiguous()
    log2s, H = float(np.log2(enc.per_level_scale)), enc.base_resolution
    out_blc = torch.empty(S, 32, dtype=torch.half, device=dev)
    out_lbc = torch.empty(16, S, 2, dtype=torch.half, device=dev)
    dummy = torch.empty(1, dtype=torch.half, device=dev)
    row("grid_encode_forward (fp16 table)", lambda: GB.grid_encode_forward(x, table, offsets, out_blc, S, 3, 2, 16, log2s, H, False, dummy, 0, 1),
        (lambda: R.grid_encode_forward(x, table, offsets, out_lbc, S, 3, 2, 16, log2s, H, False, dummy, 0)) if R else None,
        "reference writes [L,B,C]; its wrapper then copies to [B,L*C] (grid.py:52), not included")
    grad_blc = (torch.randn(S, 32, device=dev) * 1e-2).half()
    grad_lbc = grad_blc.view(S, 16, 2).permute(1, 0, 2).contiguous()
    gt32 = torch.empty(table.shape, dtype=torch.float32, device=dev)
    gt16 = torch.empty(table.shape, dtype=torch.half, device=dev)

    def scatter_ours():
        gt32.zero_()
        GB.grid_encode_backward(grad_blc, x, table, offsets, gt32, S, 3, 2, 16, log2s, H, False, dummy, dummy, 0, 1)

    def scatter_ref():
        gt16.zero_()
        R.grid_encode_backward(grad_lbc, x, table, offsets, gt16, S, 3, 2, 16, log2s, H, False, dummy, dummy, 0)

    row("grid_encode_backward (fp16 table)", scatter_ours, scatter_ref if R else None,
        "incl. zero-fill of the gradient table; reference: fp16 atomics into an fp16 table, and its wrapper first copies grad to [L,B,C] "
        "(grid.py:70), not included; ours: fp32 accumulation")
    del out_lbc, grad_lbc, gt16, gt32, grad_blc

    # ---------------- fully-fused MLP (K16-K19): sigma-net (2 layers) and colour-net (3 layers), 32 -> 64 -> ... -> 16
    R = ref.load("_ffmlp")
    xin = out_blc
    for nl, name in ((2, "sigma-net 32-64-64-16"), (3, "colour-net 32-64-64-64-16")):
        w = ((torch.rand(64 * (32 + 64 * (nl - 1) + 16), device=dev) * 2 - 1) * (3 / 64) ** 0.5).half()
        out = torch.empty(S, 16, dtype=torch.half, device=dev)
        fb = torch.empty(nl, S, 64, dtype=torch.half, device=dev)
        ib = torch.empty(S, 64, dtype=torch.half, device=dev)
        g = (torch.randn(S, 16, device=dev) * 0.1).half()
        bb = torch.empty(nl, S, 64, dtype=torch.half, device=dev)
        gi = torch.empty(S, 32, dtype=torch.half, device=dev)
        gw16 = torch.empty(w.numel(), dtype=torch.half, device=dev)
        gw32 = torch.empty(w.numel(), dtype=torch.float32, device=dev)
        if R:
            R.allocate_splitk(nl + 1)
        row(f"ffmlp_inference {name}", lambda: FB.ffmlp_inference(xin, w, S, 32, 16, 64, nl, 0, 6, None, out),
            (lambda: R.ffmlp_inference(xin, w, S, 32, 16, 64, nl, 0, 6, ib, out)) if R else None)
        row(f"ffmlp_forward (training, stores forward_buffer) {name}", lambda: FB.ffmlp_forward(xin, w, S, 32, 16, 64, nl, 0, 6, fb, out),
            (lambda: R.ffmlp_forward(xin, w, S, 32, 16, 64, nl, 0, 6, fb, out)) if R else None)
        FB.ffmlp_forward(xin, w, S, 32, 16, 64, nl, 0, 6, fb, out)
        row(f"ffmlp_backward (from forward_buffer) {name}", lambda: FB.ffmlp_backward(g, xin, w, fb, S, 32, 16, 64, nl, 0, 6, True, None, gi, gw32),
            (lambda: R.ffmlp_backward(g, xin, w, fb, S, 32, 16, 64, nl, 0, 6, True, bb, gi, gw16)) if R else None,
            "reference: dgrad kernel + (nl+1) CUTLASS split-K weight-gradient GEMMs on side streams, fp16 accumulation")
        row(f"ffmlp_backward (recomputing, no forward_buffer: the training path) {name}",
            lambda: FB.ffmlp_backward(g, xin, w, None, S, 32, 16, 64, nl, 0, 6, True, None, gi, gw32), None)
        del out, fb, ib, g, bb, gi
    both = [r for r in rows if "ours_ms" in r and "reference_build_ms" in r and "stores forward_buffer" not in r["kernel"] and "inference" not in r["kernel"]]
    res["training_step_kernels_ms"] = {"ours": round(sum(r["ours_ms"] for r in both), 4), "reference_build": round(sum(r["reference_build_ms"] for r in both), 4),
                                       "rows": [r["kernel"] for r in both]}
    try:
        res["full_step"] = full_step(n_rays, bound, iters)
    except Exception as e:  # noqa: BLE001
        res["full_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return res


def full_step(n_rays=4096, bound=3, iters=3):
    """One whole training step (fwd + bwd + GradScaler + torch Adam, fp16 autocast, 3 colour channels) of the reference's hot path —
    nerf/network_ff.py + NeRFRenderer.run_cuda chained by oracle/ref_chain.RefStack — (a) on the reference's own CUDA build and
    (b) with the same unfused wrappers on this repo's kernels (INTEGRATION.md level 2: only `_backend` swapped).  ms per step, median."""
    from . import ref_chain
    dev = torch.device("cuda", torch.cuda.current_device())
    o, d = synthetic.random_rays(n_rays, bound, seed=100)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    target = torch.rand(n_rays, 3, device=dev)
    out = {"rays": n_rays, "unit": "ms per training step (eager launches, median of %d)" % iters}
    for which in ("reference", "ours"):
        ref_chain.use_backends(which)
        try:
            torch.manual_seed(0)
            m = ref_chain.RefStack(bound=bound).to(dev).train()
            with torch.no_grad():
                m.encoder.embeddings.uniform_(-1e-4, 1e-4)
                m.w_sigma.uniform_(-(3 / 64) ** 0.5, (3 / 64) ** 0.5)
                m.w_color.uniform_(-(3 / 64) ** 0.5, (3 / 64) ** 0.5)
            grid = synthetic.ball_density_grid(bound, m.cascade)
            m.density_grid.copy_(torch.from_numpy(grid))
            m.density_bitfield.copy_(torch.from_numpy(synthetic.packbits_np(grid)))
            opt = torch.optim.Adam(m.parameters(), lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
            scaler = torch.amp.GradScaler("cuda")

            def step():
                res = m.render_train(o, d, bg_color=1, perturb=True)
                loss = ((res["image"].float() - target) ** 2).mean()
                opt.zero_grad(set_to_none=True)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()

            step()                                         # exact sizing (mean_count = 0), then fixed buffers like the trainer
            m.mean_count = int(m.step_counter[0, 0].item())
            key = "reference_build_ms" if which == "reference" else "reference_wrappers_on_this_repos_kernels_ms"
            out[key] = timeit(step, iters)
            out["samples_per_step"] = m.mean_count
            del m, opt
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[which + "_error"] = f"{type(e).__name__}: {e}"[:300]
            torch.cuda.synchronize()
        finally:
            ref_chain.use_backends("reference")
    return out

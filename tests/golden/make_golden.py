#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own CUDA extensions (oracle/_ref, built
by oracle/build_ref.py from the unmodified sources under /root/reference) on a B200.

The reference ships no golden vectors; these files are the pin: outputs of the
reference build on small seeded inputs.  tests/test_golden.py checks (CPU) that the oracle
reproduces them and (GPU) that this repo's kernels do.  Inputs are regenerated from the seeds
below (`inputs(...)`), only the reference outputs are stored.

Run on the GPU box:  python tests/golden/make_golden.py        (writes next to this file)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.environ.get("GOLDEN_OUT", HERE)      # the GPU box only returns gpurun_out/: GOLDEN_OUT=gpurun_out/golden
KEEP_RAYS = 12                                 # rays whose full sample lists are stored (counts are stored for all)
sys.path.insert(0, ROOT)

from enerf_b200 import synthetic  # noqa: E402
from oracle import oracle  # noqa: E402


# ------------------------------------------------------------------ seeded inputs (shared with the tests)
def grid_inputs(bound, n=2048):
    rng = np.random.default_rng(1000 + bound)
    pls = oracle.per_level_scale_for(2048 * bound)
    offsets = oracle.grid_offsets(3, 16, pls, 16, 19)
    emb = rng.uniform(-1, 1, size=(offsets[-1], 2)).astype(np.float32)
    x = rng.random((n, 3)).astype(np.float32)
    x[-1] = (1.0, 0.0, 0.5)
    x[-2] = (1.5, 0.5, 0.5)          # out of range -> zeros
    grad = rng.normal(size=(16, n, 2)).astype(np.float32) * 0.01
    return pls, offsets, emb, x, grad


def sh_inputs(n=1024):
    rng = np.random.default_rng(7)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return d.astype(np.float32)


def march_inputs(bound, n_rays=128):
    cascade = 1 + int(np.ceil(np.log2(bound)))
    bits = synthetic.packbits_np(synthetic.ball_density_grid(bound, cascade))
    o, d = synthetic.random_rays(n_rays, bound, seed=50 + bound)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    return cascade, bits, o, d, aabb


def composite_inputs(rays, M, n_ch=3):
    rng = np.random.default_rng(11)
    sig = (rng.uniform(0, 40, M) * (rng.random(M) < 0.7)).astype(np.float32)
    rgb = rng.random((M, n_ch)).astype(np.float32)
    g_ws = rng.normal(size=rays.shape[0]).astype(np.float32)
    g_im = rng.normal(size=(rays.shape[0], n_ch)).astype(np.float32)
    return sig, rgb, g_ws, g_im


def ffmlp_inputs(nl, B=256):
    import torch
    torch.manual_seed(42)                                       # ffmlp/ffmlp.py:141-144
    nw = 64 * (32 + 64 * (nl - 1) + 16)
    w = torch.empty(nw).uniform_(-(3 / 64) ** 0.5, (3 / 64) ** 0.5).half().numpy()
    rng = np.random.default_rng(nl)
    x = (rng.normal(size=(B, 32)) * 0.5).astype(np.float16)
    g = (rng.normal(size=(B, 16)) * 0.1).astype(np.float16)
    return w, x, g


def table_checksums(gg, offsets, seed=5):
    """per-level checksums of a gradient table [n,2] (float64): sum, sum|.|, and a seeded random projection"""
    rng = np.random.default_rng(seed)
    proj = rng.normal(size=gg.shape)
    L = len(offsets) - 1
    out = {"bwd_sum": np.zeros(L), "bwd_abs": np.zeros(L), "bwd_proj": np.zeros(L), "bwd_nnz": np.zeros(L, np.int64)}
    for l in range(L):
        a = gg[offsets[l]:offsets[l + 1]]
        out["bwd_sum"][l], out["bwd_abs"][l] = a.sum(), np.abs(a).sum()
        out["bwd_proj"][l] = (a * proj[offsets[l]:offsets[l + 1]]).sum()
        out["bwd_nnz"][l] = np.count_nonzero(np.abs(a).sum(-1))
    return out


def sort_by_ray(rays, *arrays):
    """ray-ordered concatenation (the reference's atomics make the buffer order arbitrary)"""
    order = np.argsort(rays[:, 0], kind="stable")
    outs = [[] for _ in arrays]
    counts = np.zeros(rays.shape[0], np.int32)
    for rid, off, cnt in rays[order]:
        counts[rid] = cnt
        for o, a in zip(outs, arrays):
            o.append(a[off:off + cnt])
    return counts, [np.concatenate(o) for o in outs]


def main():
    import torch

    from oracle import ref
    dev = "cuda"

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    os.makedirs(OUT, exist_ok=True)
    RM, GE, SH, FF = (ref.load(n) for n in ref.NAMES)
    assert all(m is not None for m in (RM, GE, SH, FF)), "oracle/_ref is not built"

    # ---- hash grid: forward (fp32 + fp16 table), dy_dx, backward (fp32 table)
    for bound in (1, 2, 3):
        pls, offsets, emb, x, grad = grid_inputs(bound)
        B = len(x)
        out = {}
        for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
            e = t(emb).to(dt)
            o = torch.empty(16, B, 2, device=dev, dtype=dt)
            dy = torch.empty(B, 16 * 3 * 2, device=dev, dtype=dt)
            GE.grid_encode_forward(t(x), e, t(offsets), o, B, 3, 2, 16, float(np.log2(pls)), 16, True, dy, 0)
            out["fwd_" + name] = o.cpu().numpy()
            out["dydx_" + name] = dy.cpu().numpy()[:256]
        gg = torch.zeros(int(offsets[-1]), 2, device=dev)
        dummy = torch.zeros(1, device=dev)
        GE.grid_encode_backward(t(grad), t(x), t(emb), t(offsets), gg, B, 3, 2, 16, float(np.log2(pls)), 16, False, dummy, dummy, 0)
        out.update(table_checksums(gg.cpu().numpy().astype(np.float64), offsets))
        # the device's per-level scales (exp2f differs from libm by an ulp on some levels)
        lv = torch.arange(16, device=dev, dtype=torch.float32)
        out["level_scales"] = (torch.exp2(lv * float(np.float32(np.log2(pls)))) * 16.0 - 1.0).cpu().numpy()
        np.savez_compressed(os.path.join(OUT, f"grid_bound{bound}.npz"), **out)

    # ---- SH degree 1..8 (fp32)
    d = sh_inputs()
    out = {}
    for deg in range(1, 9):
        o = torch.empty(len(d), deg * deg, device=dev)
        dy = torch.empty(len(d), 3 * deg * deg, device=dev)
        SH.sh_encode_forward(t(d), o, len(d), 3, deg, True, dy)
        out[f"deg{deg}"] = o.cpu().numpy()
        out[f"dydx{deg}"] = dy.cpu().numpy()[:128]
    np.savez_compressed(os.path.join(OUT, "sh.npz"), **out)

    # ---- marcher + compositing
    for bound in (1, 3):
        cascade, bits, o, d, aabb = march_inputs(bound)
        N = len(o)
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        RM.near_far_from_aabb(t(o), t(d), t(aabb), N, 0.2, nears, fars)
        out = {"nears": nears.cpu().numpy(), "fars": fars.cpu().numpy()}
        for tag, perturb, dt_gamma in (("plain", 0, 0.0), ("perturb", 1, 0.0), ("cone", 1, 1.0 / 128)):
            M = N * 1024
            xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
            rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
            counter = torch.zeros(2, dtype=torch.int32, device=dev)
            RM.march_rays_train(t(o), t(d), t(bits), float(bound), dt_gamma, 1024, N, cascade, 128, M, nears, fars, xyzs, dirs, deltas, rays,
                                counter, perturb)
            r = rays.cpu().numpy()
            counts, (sx, sdl) = sort_by_ray(r, xyzs.cpu().numpy(), deltas.cpu().numpy())
            keep = int(counts[:KEEP_RAYS].sum())
            out[f"{tag}_counts"] = counts
            out[f"{tag}_xyzs"] = sx[:keep]
            out[f"{tag}_deltas"] = sdl[:keep]
            out[f"{tag}_xyz_sum"] = sx.astype(np.float64).sum(0)            # checksum over ALL rays' samples
            out[f"{tag}_delta_sum"] = sdl.astype(np.float64).sum(0)
            if tag == "perturb":
                m = int(counter[0])
                sig, rgb, g_ws, g_im = composite_inputs(r, m + 1)
                ws, dp, im = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
                RM.composite_rays_train_forward(t(sig), t(rgb), deltas[:m + 1].contiguous(), rays, m + 1, N, ws, dp, im)
                gs, gr = torch.zeros(m + 1, device=dev), torch.zeros(m + 1, 3, device=dev)
                RM.composite_rays_train_backward(t(g_ws), t(g_im), t(sig), t(rgb), deltas[:m + 1].contiguous(), rays, ws, im, m + 1, N, gs, gr)
                out["comp_rays"] = r
                out["comp_ws"], out["comp_depth"], out["comp_image"] = ws.cpu().numpy(), dp.cpu().numpy(), im.cpu().numpy()
                gsn, grn = gs.cpu().numpy(), gr.cpu().numpy()
                out["comp_grad_sigmas"], out["comp_grad_rgbs"] = gsn[:4096], grn[:4096]
                out["comp_grad_sums"] = np.array([gsn.astype(np.float64).sum(), np.abs(gsn).astype(np.float64).sum(), grn.astype(np.float64).sum()])
                out["comp_deltas"] = deltas[:m + 1].cpu().numpy()
        np.savez_compressed(os.path.join(OUT, f"march_bound{bound}.npz"), **out)

    # ---- FFMLP (reference: fp16 accumulation)
    out = {}
    for nl in (2, 3):
        w, x, g = ffmlp_inputs(nl)
        B = len(x)
        FF.allocate_splitk(nl + 1)
        o = torch.empty(B, 16, device=dev, dtype=torch.half)
        fb = torch.empty(nl, B, 64, device=dev, dtype=torch.half)
        FF.ffmlp_forward(t(x), t(w), B, 32, 16, 64, nl, 0, 6, fb, o)
        bb = torch.zeros(nl, B, 64, device=dev, dtype=torch.half)
        gi = torch.zeros(B, 32, device=dev, dtype=torch.half)
        gw = torch.zeros(len(w), device=dev, dtype=torch.half)
        FF.ffmlp_backward(t(g), t(x), t(w), fb, B, 32, 16, 64, nl, 0, 6, True, bb, gi, gw)
        torch.cuda.synchronize()
        out[f"out{nl}"], out[f"fb{nl}"] = o.cpu().numpy(), fb.cpu().numpy()
        out[f"bb{nl}"], out[f"gi{nl}"], out[f"gw{nl}"] = bb.cpu().numpy(), gi.cpu().numpy(), gw.cpu().numpy()
    np.savez_compressed(os.path.join(OUT, "ffmlp.npz"), **out)
    print("golden vectors written to", OUT, sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()

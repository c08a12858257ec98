"""Golden vectors = outputs of the REFERENCE's own CUDA build on a B200 (tests/golden/make_golden.py).

CPU tests: the oracle reproduces them (this is what pins the oracle to the reference).
GPU tests: this repo's kernels reproduce them (parity with the reference on identical inputs).
Bit-exact for the marcher / near-far / hash-grid forward; stated tolerances elsewhere."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.golden import make_golden as mg

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def _ulp_close(got, want, dtype):
    same = (got == want).mean()
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(dtype)).astype(np.float64)
    return same, float((diff / ulp).max())


# ============================================================================ CPU: oracle vs golden
@pytest.mark.parametrize("bound", [1, 2, 3])
def test_oracle_grid_forward_reproduces_reference(bound):
    g = load(f"grid_bound{bound}.npz")
    pls, offsets, emb, x, grad = mg.grid_inputs(bound)
    for name, dt in (("f32", np.float32), ("f16", np.float16)):
        out, dy = oracle.grid_encode_forward(x, emb.astype(dt), offsets, pls, 16, calc_grad_inputs=True, level_scales=g["level_scales"])
        same, ulps = _ulp_close(out, g["fwd_" + name], dt)
        assert same >= 0.999 and ulps <= 1.0, (name, same, ulps)
        same, ulps = _ulp_close(dy.reshape(len(x), -1)[:256], g["dydx_" + name], dt)
        assert same >= 0.995 and ulps <= 2.0, ("dydx " + name, same, ulps)
    assert np.all(g["fwd_f32"][:, -2] == 0)          # the out-of-range point


@pytest.mark.parametrize("bound", [1, 2, 3])
def test_oracle_grid_backward_reproduces_reference_checksums(bound):
    g = load(f"grid_bound{bound}.npz")
    pls, offsets, emb, x, grad = mg.grid_inputs(bound)
    gg = oracle.grid_encode_backward(grad, x, offsets, offsets[-1], 2, pls, 16, level_scales=g["level_scales"])
    chk = mg.table_checksums(gg, offsets)
    assert np.array_equal(chk["bwd_nnz"], g["bwd_nnz"])
    for k in ("bwd_sum", "bwd_abs", "bwd_proj"):
        assert np.allclose(chk[k], g[k], rtol=2e-4, atol=2e-4 * np.abs(g["bwd_abs"]).max()), k


def test_oracle_sh_reproduces_reference():
    g = load("sh.npz")
    d = mg.sh_inputs()
    for deg in range(1, 9):
        want = g[f"deg{deg}"]
        got = oracle.sh_encode(d, deg)
        assert np.abs(got - want).max() < (2e-6 if deg <= 4 else 3e-5), deg


@pytest.mark.parametrize("bound", [1, 3])
def test_oracle_marcher_and_composite_reproduce_reference(bound):
    g = load(f"march_bound{bound}.npz")
    cascade, bits, o, d, aabb = mg.march_inputs(bound)
    nears, fars = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    assert np.array_equal(nears, g["nears"]) and np.array_equal(fars, g["fars"])
    for tag, perturb, dt_gamma in (("plain", False, 0.0), ("perturb", True, 0.0), ("cone", True, 1.0 / 128)):
        xyzs, dirs, deltas, rays, cnt = oracle.march_rays_train(o, d, bound, bits, cascade, 128, nears, fars, perturb=perturb, dt_gamma=dt_gamma)
        counts, (sx, sdl) = mg.sort_by_ray(rays, xyzs, deltas)
        assert np.array_equal(counts, g[f"{tag}_counts"]), tag
        keep = int(counts[:mg.KEEP_RAYS].sum())
        assert np.array_equal(sx[:keep], g[f"{tag}_xyzs"]) and np.array_equal(sdl[:keep], g[f"{tag}_deltas"]), tag
        assert np.allclose(sx.astype(np.float64).sum(0), g[f"{tag}_xyz_sum"], rtol=1e-12, atol=1e-9)
        assert np.allclose(sdl.astype(np.float64).sum(0), g[f"{tag}_delta_sum"], rtol=1e-12, atol=1e-9)
    r = g["comp_rays"]
    m1 = g["comp_deltas"].shape[0]
    sig, rgb, g_ws, g_im = mg.composite_inputs(r, m1)
    ws, dp, im = oracle.composite_rays_train_forward(sig, rgb, g["comp_deltas"], r)
    assert np.allclose(ws, g["comp_ws"], atol=1e-5) and np.allclose(dp, g["comp_depth"], atol=1e-5) and np.allclose(im, g["comp_image"], atol=1e-5)
    gs, gr = oracle.composite_rays_train_backward(g_ws, g_im, sig, rgb, g["comp_deltas"], r, g["comp_ws"], g["comp_image"])
    assert np.allclose(gs[:4096], g["comp_grad_sigmas"], atol=2e-5, rtol=1e-3) and np.allclose(gr[:4096], g["comp_grad_rgbs"], atol=1e-6, rtol=1e-4)


def test_oracle_ffmlp_is_consistent_with_reference_fp16_kernels():
    g = load("ffmlp.npz")
    for nl in (2, 3):
        w, x, gy = mg.ffmlp_inputs(nl)
        y, fb = oracle.ffmlp_forward(x, w, 32, 64, nl)
        scale = np.abs(y).max()
        assert np.abs(g[f"out{nl}"].astype(np.float64) - y).max() < 3e-2 * scale
        assert np.abs(g[f"fb{nl}"].astype(np.float64) - fb.astype(np.float64)).max() < 3e-2 * np.abs(fb.astype(np.float64)).max()
        gx, gw, bb = oracle.ffmlp_backward(gy, x, w, g[f"fb{nl}"], 32, 64, nl)
        assert np.abs(g[f"gi{nl}"].astype(np.float64) - gx).max() < 3e-2 * np.abs(gx).max()
        assert np.abs(g[f"gw{nl}"].astype(np.float64) - gw).max() < 5e-2 * np.abs(gw).max()


# ============================================================================ GPU: this repo vs golden
def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("bound", [1, 2, 3])
def test_gpu_grid_matches_golden(bound):
    from enerf_b200.backends import gridencoder_backend as GB
    g = load(f"grid_bound{bound}.npz")
    pls, offsets, emb, x, grad = mg.grid_inputs(bound)
    B = len(x)
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        e = _t(emb).to(dt)
        out = torch.empty(16, B, 2, device="cuda", dtype=dt)
        dy = torch.empty(B, 96, device="cuda", dtype=dt)
        GB.grid_encode_forward(_t(x), e, _t(offsets), out, B, 3, 2, 16, np.log2(pls), 16, True, dy, 0, 0)
        assert np.array_equal(out.cpu().numpy(), g["fwd_" + name]), name            # bit-exact vs the reference build
        assert np.array_equal(dy.cpu().numpy()[:256], g["dydx_" + name]), name
    gg = torch.zeros(int(offsets[-1]), 2, device="cuda")
    dummy = torch.zeros(1, device="cuda")
    GB.grid_encode_backward(_t(grad), _t(x), _t(emb), _t(offsets), gg, B, 3, 2, 16, np.log2(pls), 16, False, dummy, dummy, 0, 0)
    chk = mg.table_checksums(gg.cpu().numpy().astype(np.float64), offsets)
    assert np.array_equal(chk["bwd_nnz"], g["bwd_nnz"])
    for k in ("bwd_sum", "bwd_abs", "bwd_proj"):
        assert np.allclose(chk[k], g[k], rtol=2e-4, atol=2e-4 * np.abs(g["bwd_abs"]).max()), k


@pytest.mark.gpu
def test_gpu_sh_matches_golden():
    from enerf_b200.backends import shencoder_backend as SB
    g = load("sh.npz")
    d = mg.sh_inputs()
    for deg in range(1, 9):
        out = torch.empty(len(d), deg * deg, device="cuda")
        dy = torch.empty(len(d), 3 * deg * deg, device="cuda")
        SB.sh_encode_forward(_t(d), out, len(d), 3, deg, True, dy)
        assert np.abs(out.cpu().numpy() - g[f"deg{deg}"]).max() < 3e-5, deg
        assert np.abs(dy.cpu().numpy()[:128] - g[f"dydx{deg}"]).max() < 3e-4, deg


@pytest.mark.gpu
@pytest.mark.parametrize("bound", [1, 3])
def test_gpu_marcher_and_composite_match_golden(bound):
    from enerf_b200 import raymarching as rm
    g = load(f"march_bound{bound}.npz")
    cascade, bits, o, d, aabb = mg.march_inputs(bound)
    nears, fars = rm.near_far_from_aabb(_t(o), _t(d), _t(aabb), 0.2)
    assert np.array_equal(nears.cpu().numpy(), g["nears"]) and np.array_equal(fars.cpu().numpy(), g["fars"])
    for tag, perturb, dt_gamma in (("plain", False, 0.0), ("perturb", True, 0.0), ("cone", True, 1.0 / 128)):
        counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        xyzs, dirs, deltas, rays = rm.march_rays_train(_t(o), _t(d), float(bound), _t(bits), cascade, 128, nears, fars, counter, -1, perturb, 128,
                                                       False, dt_gamma, 1024)
        counts, (sx, sdl) = mg.sort_by_ray(rays.cpu().numpy(), xyzs.cpu().numpy(), deltas.cpu().numpy())
        assert np.array_equal(counts, g[f"{tag}_counts"]), tag
        keep = int(counts[:mg.KEEP_RAYS].sum())
        assert np.array_equal(sx[:keep], g[f"{tag}_xyzs"]) and np.array_equal(sdl[:keep], g[f"{tag}_deltas"]), tag
        assert np.allclose(sx.astype(np.float64).sum(0), g[f"{tag}_xyz_sum"], rtol=1e-12, atol=1e-9)
    r = g["comp_rays"]
    m1 = g["comp_deltas"].shape[0]
    sig, rgb, g_ws, g_im = mg.composite_inputs(r, m1)
    ts, tr = _t(sig).requires_grad_(True), _t(rgb).requires_grad_(True)
    ws, dp, im = rm.composite_rays_train(ts, tr, _t(g["comp_deltas"]), _t(r))
    assert np.allclose(ws.detach().cpu().numpy(), g["comp_ws"], atol=1e-5) and np.allclose(im.detach().cpu().numpy(), g["comp_image"], atol=1e-5)
    assert np.allclose(dp.detach().cpu().numpy(), g["comp_depth"], atol=1e-5)
    ((ws * _t(g_ws)).sum() + (im * _t(g_im)).sum()).backward()
    assert np.allclose(ts.grad.cpu().numpy()[:4096], g["comp_grad_sigmas"], atol=2e-5, rtol=1e-3)
    assert np.allclose(tr.grad.cpu().numpy()[:4096], g["comp_grad_rgbs"], atol=1e-6, rtol=1e-4)


@pytest.mark.gpu
def test_gpu_ffmlp_matches_golden_within_fp16_accumulation_error():
    from enerf_b200.backends import ffmlp_backend as FB
    g = load("ffmlp.npz")
    for nl in (2, 3):
        w, x, gy = mg.ffmlp_inputs(nl)
        B = len(x)
        out = torch.empty(B, 16, device="cuda", dtype=torch.half)
        fb = torch.empty(nl, B, 64, device="cuda", dtype=torch.half)
        FB.ffmlp_forward(_t(x), _t(w), B, 32, 16, 64, nl, 0, 6, fb, out)
        y, _ = oracle.ffmlp_forward(x, w, 32, 64, nl)
        ours, ref = out.cpu().numpy().astype(np.float64), g[f"out{nl}"].astype(np.float64)
        assert np.abs(ours - ref).max() < 3e-2 * np.abs(ref).max()
        assert np.abs(ours - y).max() <= np.abs(ref - y).max() + 1e-3          # at least as close to exact as the reference
        gi = torch.empty(B, 32, device="cuda", dtype=torch.half)
        gw = torch.empty(len(w), device="cuda", dtype=torch.float32)
        FB.ffmlp_backward(_t(gy), _t(x), _t(w), _t(g[f"fb{nl}"]), B, 32, 16, 64, nl, 0, 6, True, None, gi, gw)
        assert np.abs(gi.cpu().numpy().astype(np.float64) - g[f"gi{nl}"].astype(np.float64)).max() < 3e-2 * np.abs(g[f"gi{nl}"].astype(np.float64)).max()
        assert np.abs(gw.cpu().numpy().astype(np.float64) - g[f"gw{nl}"].astype(np.float64)).max() < 5e-2 * np.abs(g[f"gw{nl}"].astype(np.float64)).max()

"""GPU parity at BASELINE.json's full sizes (configs[1]: 4096 rays, bound 3 -> ~3.3 M samples; configs[4]: 65 536 rays -> ~50 M samples;
configs[3]: one 800 x 800 frame, 640 000 rays, through the inference loop).

The oracle needs minutes at this size, so the kernels are checked through properties that do not
depend on the size and tie the full-size launch to the small launches the oracle *does* verify
(tests/test_gpu_encoders.py, tests/test_gpu_ffmlp.py):

  * row independence — a sample's result in the full launch equals its result in a small launch
    of a random subset (bit-exact: same instruction sequence per sample/row);
  * partition of unity — a table of ones encodes to ones (the eight blend weights sum to 1);
  * exact homogeneity — scaling an fp32 table by 4 scales the encoding by exactly 4;
  * checksum of checksums — per level, the sum of the scattered table gradient equals the sum of
    the incoming gradient column (again: weights sum to 1);
  * additivity — the gradient of the whole batch equals the sum of the gradients of its halves.

Tolerances: bit-exact where stated; sums of fp32 atomics in arbitrary order are compared in
float64 with 1e-5 relative (checksums) / 1e-4 of the largest entry (additivity); fp16 MLP outputs
against a float32 torch evaluation of the same network: 2e-3 * max|value| (tests/test_gpu_ffmlp.py).
"""
import numpy as np
import pytest
import torch

from enerf_b200 import raymarching as rm
from enerf_b200.backends import ffmlp_backend as FB
from enerf_b200.backends import gridencoder_backend as GB
from oracle import oracle
from tests.gpu_common import DEV, scene, t

pytestmark = pytest.mark.gpu

BOUND, N_RAYS = 3, 4096


@pytest.fixture(scope="module")
def samples():
    """unit-cube coordinates of every sample the marcher emits for the BASELINE batch, marcher order"""
    sc = scene(N_RAYS, BOUND, seed=11)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, _, _, _ = rm.march_rays_train(t(sc["o"]), t(sc["d"]), float(BOUND), t(sc["bits"]), 3, 128, t(sc["nears"]), t(sc["fars"]), counter,
                                        -1, True, 128, False, 0, 1024)
    m = int(counter[0])
    assert m > 2_000_000, m
    return ((xyzs[:m] + BOUND) / (2 * BOUND)).contiguous()


def _grid():
    pls = oracle.per_level_scale_for(2048 * BOUND, 16, 16)
    offsets = oracle.grid_offsets(3, 16, pls, 16, 19)
    return pls, t(offsets), int(offsets[-1]), offsets


def _encode(x, emb, toff, pls):
    B = x.shape[0]
    out = torch.empty(B, 32, device=DEV, dtype=emb.dtype)
    dummy = torch.empty(1, device=DEV, dtype=emb.dtype)
    GB.grid_encode_forward(x, emb, toff, out, B, 3, 2, 16, np.log2(pls), 16, False, dummy, 0, 1)
    return out


def _scatter(grad, x, emb, toff, n_entries, pls):
    B = x.shape[0]
    gg = torch.zeros(n_entries, 2, device=DEV, dtype=torch.float32)
    dummy = torch.zeros(1, device=DEV, dtype=grad.dtype)
    GB.grid_encode_backward(grad, x, emb, toff, gg, B, 3, 2, 16, np.log2(pls), 16, False, dummy, dummy, 0, 1)
    return gg


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_gather_rows_are_independent_of_the_launch_size(samples, dtype):
    pls, toff, n_entries, _ = _grid()
    g = torch.Generator(device=DEV).manual_seed(1)
    emb = (torch.rand(n_entries, 2, device=DEV, generator=g) * 2 - 1).to(dtype)
    full = _encode(samples, emb, toff, pls)
    pick = torch.randperm(samples.shape[0], device=DEV, generator=g)[:8191]          # ragged on purpose
    small = _encode(samples[pick].contiguous(), emb, toff, pls)
    assert torch.equal(full[pick], small)
    assert torch.isfinite(full.float()).all()


def test_gather_partition_of_unity_and_homogeneity(samples):
    pls, toff, n_entries, _ = _grid()
    ones = torch.ones(n_entries, 2, device=DEV)
    enc = _encode(samples, ones, toff, pls)
    assert float((enc - 1).abs().max()) < 1e-5
    g = torch.Generator(device=DEV).manual_seed(2)
    emb = torch.rand(n_entries, 2, device=DEV, generator=g) * 2 - 1
    assert torch.equal(_encode(samples, emb * 4, toff, pls), _encode(samples, emb, toff, pls) * 4)


def test_scatter_checksums_and_additivity(samples):
    pls, toff, n_entries, offsets = _grid()
    B = samples.shape[0]
    g = torch.Generator(device=DEV).manual_seed(3)
    emb = torch.zeros(n_entries, 2, device=DEV, dtype=torch.float16)
    grad = (torch.randn(B, 32, device=DEV, generator=g).abs() * 1e-2).half()
    gg = _scatter(grad, samples, emb, toff, n_entries, pls)
    col = grad.double().sum(0).view(16, 2)
    for lv in range(16):
        got = gg[offsets[lv]:offsets[lv + 1]].double().sum(0)
        assert torch.allclose(got, col[lv], rtol=1e-5, atol=0), (lv, got.tolist(), col[lv].tolist())
    h = (B // 2) // 128 * 128 + 37                                                   # ragged split
    parts = _scatter(grad[:h].contiguous(), samples[:h].contiguous(), emb, toff, n_entries, pls) + \
        _scatter(grad[h:].contiguous(), samples[h:].contiguous(), emb, toff, n_entries, pls)
    assert float((gg - parts).abs().max()) <= 1e-4 * float(gg.abs().max())


def _net(I, W, nl, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    nw = W * (I + W * (nl - 1) + 16)
    return ((torch.rand(nw, device=DEV, generator=g) * 2 - 1) * (3 / W) ** 0.5).half()


def _torch_mlp(x, w, I, W, nl):
    """float32 evaluation with fp16 storage of the hidden activations (what forward_buffer holds)"""
    h, pos = x.float(), 0
    for layer in range(nl + 1):
        rows = 16 if layer == nl else W
        cols = I if layer == 0 else W
        m = w[pos:pos + rows * cols].view(rows, cols).float()
        pos += rows * cols
        h = h @ m.t()
        if layer < nl:
            h = torch.relu(h).half().float()
    return h


@pytest.mark.parametrize("nl", [2, 3])
def test_mlp_rows_are_independent_of_the_launch_size(samples, nl):
    I, W = 32, 64
    B = samples.shape[0] // 128 * 128
    g = torch.Generator(device=DEV).manual_seed(4 + nl)
    x = (torch.randn(B, I, device=DEV, generator=g) * 0.5).half()
    w = _net(I, W, nl, seed=nl)
    out = torch.empty(B, 16, device=DEV, dtype=torch.half)
    FB.ffmlp_inference(x, w, B, I, 16, W, nl, 0, 6, None, out)
    want = _torch_mlp(x, w, I, W, nl)
    assert float((out.float() - want).abs().max()) <= 2e-3 * float(want.abs().max()) + 1e-3
    lo = 128 * 1000
    small = torch.empty(128 * 33, 16, device=DEV, dtype=torch.half)
    FB.ffmlp_inference(x[lo:lo + 128 * 33].contiguous(), w, 128 * 33, I, 16, W, nl, 0, 6, None, small)
    assert torch.equal(out[lo:lo + 128 * 33], small)

    # backward without a forward_buffer (the training path): input gradients are per row, weight gradients add up
    grad = (torch.randn(B, 16, device=DEV, generator=g) * 0.1).half()

    def bwd(xs, gs):
        b = xs.shape[0]
        gin = torch.empty(b, I, device=DEV, dtype=torch.half)
        gw = torch.zeros(w.numel(), device=DEV, dtype=torch.float32)
        FB.ffmlp_backward(gs, xs, w, None, b, I, 16, W, nl, 0, 6, True, None, gin, gw)
        return gin, gw

    gin, gw = bwd(x, grad)
    h = (B // 2) // 128 * 128
    gin_a, gw_a = bwd(x[:h].contiguous(), grad[:h].contiguous())
    gin_b, gw_b = bwd(x[h:].contiguous(), grad[h:].contiguous())
    assert torch.equal(gin[:h], gin_a) and torch.equal(gin[h:], gin_b)
    assert float((gw - (gw_a + gw_b)).abs().max()) <= 1e-4 * float(gw.abs().max())
    # against autograd through the float32 evaluation: ReLU units within rounding of zero may flip -> relative L2
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    h32, pos = xr, 0
    for layer in range(nl + 1):
        rows = 16 if layer == nl else W
        cols = I if layer == 0 else W
        m = wr[pos:pos + rows * cols].view(rows, cols)
        pos += rows * cols
        h32 = h32 @ m.t()
        if layer < nl:
            h32 = torch.relu(h32)
    h32.backward(grad.float())
    rel_w = float((gw - wr.grad).norm() / wr.grad.norm())
    rel_x = float((gin.float() - xr.grad).norm() / xr.grad.norm())
    assert rel_w < 2e-2 and rel_x < 2e-2, (rel_w, rel_x)


# ---- BASELINE configs[4]: 65 536 rays per batch (spiral1-shaped, bound 3) ---------------------------------------------------------
def _march(o, d, bits, nears, fars):
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, float(BOUND), bits, 3, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
    return xyzs, deltas, rays, counter


def _synthetic_field(xyzs):
    """a deterministic per-sample density / colour (so that a ray's samples carry the same values wherever the marcher put them)"""
    sigma = (torch.sin(xyzs[:, 0] * 7.0) + torch.cos(xyzs[:, 1] * 5.0) + 2.1) * 3.0
    rgb = torch.sigmoid(xyzs * 2.0)
    return sigma.contiguous(), rgb.contiguous()


def test_march_and_composite_65536_rays_match_4096_ray_launches():
    """the marcher reserves sample ranges with atomics, so the LAYOUT of a 65 536-ray launch is not reproducible — but every ray's
    sample count, samples and composited pixel must equal what the same ray gets in a 4096-ray launch (bit-exact: per-ray code paths
    do not depend on the launch size), the ranges must tile [0, total) and the counter must equal their sum"""
    N = 65536
    sc = scene(N, BOUND, seed=21)
    o, d, bits, nears, fars = (t(sc[k]) for k in ("o", "d", "bits", "nears", "fars"))
    xyzs, deltas, rays, counter = _march(o, d, bits, nears, fars)
    total = int(counter[0])
    assert int(counter[1]) == N and total == int(rays[:, 2].sum()) and total > 30_000_000
    order = torch.argsort(rays[:, 1])
    off, cnt = rays[order, 1].long(), rays[order, 2].long()
    assert int(off[0]) == 0 and bool((off[1:] == off[:-1] + cnt[:-1]).all())          # the ranges tile [0, total)
    sigma, rgb = _synthetic_field(xyzs)
    ws, depth, image = rm.composite_rays_train(sigma, rgb, deltas, rays)
    assert bool(torch.isfinite(image).all()) and float(ws.max()) <= 1.0 + 1e-5 and float(ws.min()) >= 0.0
    steps_full = torch.empty(N, dtype=torch.int64, device=DEV)
    steps_full[rays[:, 0].long()] = rays[:, 2].long()
    first = torch.empty(N, dtype=torch.int64, device=DEV)
    first[rays[:, 0].long()] = rays[:, 1].long()
    for c in (0, 7, 15):                                                             # three of the sixteen 4096-ray chunks
        s = slice(c * 4096, (c + 1) * 4096)
        xs, ds, rs, cn = _march(o[s].contiguous(), d[s].contiguous(), bits, nears[s].contiguous(), fars[s].contiguous())
        # NB the training jitter is seeded by the ray's index IN ITS LAUNCH (raymarching.cu:349-352), so only chunk 0 shares it
        sg, rg = _synthetic_field(xs)
        w2, d2, im2 = rm.composite_rays_train(sg, rg, ds, rs)
        if c == 0:
            small = torch.empty(4096, dtype=torch.int64, device=DEV)
            small[rs[:, 0].long()] = rs[:, 2].long()
            assert torch.equal(small, steps_full[s])
            assert torch.equal(im2, image[s]) and torch.equal(w2, ws[s]) and torch.equal(d2, depth[s])
            # and the samples themselves, ray by ray
            f2 = torch.empty(4096, dtype=torch.int64, device=DEV)
            f2[rs[:, 0].long()] = rs[:, 1].long()
            for ray in (0, 1, 2047, 4095):
                n_ = int(small[ray])
                a, b = int(first[ray]), int(f2[ray])
                assert torch.equal(xyzs[a:a + n_], xs[b:b + n_]) and torch.equal(deltas[a:a + n_], ds[b:b + n_])
        else:
            assert int(cn[0]) == int(rs[:, 2].sum()) and bool(torch.isfinite(im2).all())


# ---- BASELINE configs[3]: 800 x 800 full-frame inference -----------------------------------------------------------------------
def test_full_frame_inference_equals_its_row_bands():
    """640 000 rays through the device-counted marching loop in one call vs the same frame rendered as four bands of 200 rows (what
    four GPUs would each render): the per-ray sample sequence does not depend on which rays share a launch; n_alive, n_step and the compaction
    order all differ, and with them the points where a ray's transmittance is re-derived from its weight sum (raymarching.cu:842-899:
    T = 1 - weights_sum at the start of every round), so the images agree to rounding (1e-5 of the [0, 1] range), not to the bit"""
    from enerf_b200 import synthetic
    from enerf_b200.nerf.network_ff import NeRFNetwork
    torch.manual_seed(3)
    model = NeRFNetwork(encoding="hashgrid", bound=BOUND, cuda_ray=True, out_dim_color=1).to(DEV).eval()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.5, 0.5)
        grid = synthetic.ball_density_grid(BOUND, model.cascade)
        model.density_grid.copy_(t(grid).view(model.cascade, -1))
        model.density_bitfield.copy_(t(synthetic.packbits_np(grid)))
    pose = synthetic.look_at_poses(1, 0.6 * BOUND, seed=5)[0]
    o, d = synthetic.pinhole_rays(pose, 800, 800)
    o, d = t(o), t(d)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        full = model.render(o[None], d[None], staged=False, bg_color=1, perturb=False, out_dim_color=1)
        stats = dict(model.last_render_stats)
        bands = [model.render(o[i:i + 160000][None], d[i:i + 160000][None], staged=False, bg_color=1, perturb=False, out_dim_color=1)
                 for i in range(0, 640000, 160000)]
    img = full["image"].reshape(-1)
    assert img.shape[0] == 640000 and bool(torch.isfinite(img).all()) and stats["samples"] > 50_000_000, stats
    assert float((img - torch.cat([b["image"].reshape(-1) for b in bands])).abs().max()) <= 1e-5
    assert float((full["depth"].reshape(-1) - torch.cat([b["depth"].reshape(-1) for b in bands])).abs().max()) <= 1e-4
    assert float(img.std()) > 1e-3 and float(img.min()) >= 0.0 and float(img.max()) <= 1.0 + 1e-5          # a picture, not a constant

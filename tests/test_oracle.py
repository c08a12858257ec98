"""Pins the CPU oracle (oracle/) — the checker every CUDA parity test relies on.

The reference ships no tests or golden vectors for this path, so the oracle is
pinned here against published known answers and closed forms, and against independent
re-derivations written in numpy / torch-float64.  (Its second pin, the reference's own CUDA
build run on a B200, lives in tests/test_ref_parity.py and tests/golden/.)
"""
import numpy as np
import pytest
import torch

from enerf_b200 import synthetic
from oracle import oracle

FLT_MAX = np.float32(3.4028234663852886e38)


# ------------------------------------------------------------------------------ pcg32 / morton
def test_pcg32_published_known_answer():
    # pcg32-demo (pcg-random.org, "pcg32_srandom(42u, 54u)") first six outputs
    want = [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]
    assert oracle.pcg32_stream(42, 54, 6).tolist() == want


def test_pcg32_float_in_unit_interval():
    vals = [oracle.pcg32_first_float(n) for n in range(256)]
    assert all(0.0 <= v < 1.0 for v in vals)
    assert len(set(vals)) > 250
    # next_float = bits((u >> 9) | 0x3f800000) - 1   (pcg32.h:107-116)
    u = int(oracle.pcg32_stream(7, 1, 1)[0])
    want = np.array([(u >> 9) | 0x3f800000], dtype=np.uint32).view(np.float32)[0] - np.float32(1.0)
    assert oracle.pcg32_first_float(7) == float(want)


def test_morton_known_values_and_bijection():
    assert oracle.morton3D(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 0, 0], [127, 127, 127]])).tolist() == [1, 2, 4, 9, 2097151]
    idx = np.arange(128 ** 3, dtype=np.int32)
    coords = oracle.morton3D_invert(idx)
    assert coords.min() == 0 and coords.max() == 127
    assert np.array_equal(oracle.morton3D(coords), idx)
    assert np.array_equal(coords, np.stack([synthetic._compact3(idx.astype(np.uint32) >> k) for k in range(3)], -1).astype(np.int32))


def test_packbits_bit_order():
    rng = np.random.default_rng(0)
    grid = rng.random(128 * 8).astype(np.float32)
    got = oracle.packbits(grid, 0.5)
    want = np.packbits((grid > 0.5).reshape(-1, 8), axis=-1, bitorder="little").reshape(-1)
    assert np.array_equal(got, want)
    assert np.array_equal(synthetic.packbits_np(grid, 0.5), want)


# ------------------------------------------------------------------------------ near / far
def test_near_far_closed_form():
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    o = np.array([[-2, 0, 0], [-2, 0, 0], [0, 0, 0], [-2, 5, 0]], np.float32)
    d = np.array([[1, 1e-9, 1e-9], [1, 1e-9, 1e-9], [0.6, 0.0, 0.8], [1, 1e-9, 1e-9]], np.float32)
    nears, fars = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    assert nears[0] == pytest.approx(1.0) and fars[0] == pytest.approx(3.0)
    assert nears[2] == pytest.approx(0.2) and fars[2] == pytest.approx(1.25)     # origin inside: clamp to min_near
    assert nears[3] == FLT_MAX and fars[3] == FLT_MAX                           # misses the box
    nears2, _ = oracle.near_far_from_aabb(o[:1], d[:1], aabb, 1.5)
    assert nears2[0] == pytest.approx(1.5)


# ------------------------------------------------------------------------------ hash grid
def _np_grid_encode(x, emb, offsets, pls, H, gridtype=0):
    """independent vectorised float64 restatement (index math in uint32)"""
    B, D = x.shape
    L = len(offsets) - 1
    C = emb.shape[1]
    out = np.zeros((L, B, C))
    S = np.float32(np.log2(pls))
    primes = np.array([1, 2654435761, 805459861], dtype=np.uint32)
    for l in range(L):
        hs = np.uint32(offsets[l + 1] - offsets[l])
        scale = np.float32(np.exp2(np.float32(l) * S) * np.float32(H) - np.float32(1.0))
        res = np.uint32(np.ceil(scale)) + np.uint32(1)
        pos = (x.astype(np.float64) * np.float64(scale) + 0.5).astype(np.float32)      # single rounding, like fmaf
        pg = np.floor(pos).astype(np.uint32)
        fr = (pos - pg).astype(np.float64)
        dense = int(res + 1) ** D <= int(hs)
        for idx in range(1 << D):
            w = np.ones(B)
            pl = pg.copy()
            for d in range(D):
                if idx & (1 << d):
                    w *= fr[:, d]
                    pl[:, d] += 1
                else:
                    w *= 1 - fr[:, d]
            if dense or gridtype == 1:
                index = np.zeros(B, dtype=np.uint32)
                stride = np.uint32(1)
                for d in range(D):
                    if stride <= hs:
                        index = index + pl[:, d] * stride
                        stride = np.uint32(int(stride) * int(res + 1) & 0xffffffff)
            else:
                index = np.zeros(B, dtype=np.uint32)
                for d in range(D):
                    index ^= pl[:, d] * primes[d]
            index = index % hs
            out[l] += w[:, None] * emb[offsets[l] + index].astype(np.float64)
    return out


@pytest.mark.parametrize("bound", [1, 2, 3])
def test_grid_forward_fp32_matches_independent_numpy(bound):
    rng = np.random.default_rng(bound)
    pls = oracle.per_level_scale_for(2048 * bound)
    offsets = oracle.grid_offsets(3, 16, pls, 16, 19)
    emb = rng.uniform(-1, 1, size=(offsets[-1], 2)).astype(np.float32)
    x = rng.random((2048, 3)).astype(np.float32)
    # same per-level scales on both sides (libm's exp2f and numpy's may differ by an ulp)
    S = np.float32(np.log2(pls))
    scales = np.array([np.float32(np.exp2(np.float32(l) * S) * np.float32(16) - np.float32(1.0)) for l in range(16)], np.float32)
    got, _ = oracle.grid_encode_forward(x, emb, offsets, pls, 16, level_scales=scales)
    with np.errstate(over="ignore"):
        want = _np_grid_encode(x, emb, offsets, pls, 16)
    assert np.abs(got - want).max() < 2e-6
    free, _ = oracle.grid_encode_forward(x, emb, offsets, pls, 16)
    assert np.abs(free - want).max() < 2e-4
    assert got.shape == (16, 2048, 2)


def test_grid_offsets_match_survey():
    # totals 6 119 864 / 6 328 848 / 6 507 840 entries for bound 1 / 2 / 3
    for bound, total in [(1, 6119864), (2, 6328848), (3, 6507840)]:
        off = oracle.grid_offsets(3, 16, oracle.per_level_scale_for(2048 * bound), 16, 19)
        assert off[-1] == total
        assert off[1] == 4920            # level 0: ceil8(17^3) dense
        assert all((off[i + 1] - off[i]) == 2 ** 19 for i in range(5, 16))


def test_grid_level0_index_is_x_17y_289z():
    # one-hot table: entry e holds value e; a point in the interior of cell (3,5,7) at level 0
    offsets = oracle.grid_offsets(3, 1, 2.0, 16, 19)
    emb = np.arange(offsets[-1], dtype=np.float32)[:, None].repeat(2, 1)
    g = np.array([3, 5, 7])
    x = ((g + 0.25 - 0.5) / 15.0).astype(np.float32)[None]          # pos = g + 0.25
    got, _ = oracle.grid_encode_forward(x, emb, offsets, 2.0, 16)
    corners = [(3 + a) + 17 * (5 + b) + 289 * (7 + c) for c in (0, 1) for b in (0, 1) for a in (0, 1)]
    w = [(0.25 if a else 0.75) * (0.25 if b else 0.75) * (0.25 if c else 0.75) for c in (0, 1) for b in (0, 1) for a in (0, 1)]
    assert got[0, 0, 0] == pytest.approx(float(np.dot(corners, w)), rel=1e-5)


def test_grid_out_of_range_inputs_give_zero():
    offsets = oracle.grid_offsets(3, 4, 2.0, 16, 19)
    emb = np.ones((offsets[-1], 2), np.float32)
    x = np.array([[0.5, 0.5, 1.0001], [-1e-6, 0.2, 0.2], [0.5, 0.5, 0.5]], np.float32)
    got, dy = oracle.grid_encode_forward(x, emb, offsets, 2.0, 16, calc_grad_inputs=True)
    assert np.all(got[:, 0] == 0) and np.all(got[:, 1] == 0) and np.allclose(got[:, 2], 1.0)
    assert np.all(dy[:2] == 0)


def test_grid_fp16_rounds_after_every_corner():
    rng = np.random.default_rng(5)
    pls = oracle.per_level_scale_for(2048)
    offsets = oracle.grid_offsets(3, 16, pls, 16, 19)
    emb = rng.uniform(-1, 1, size=(offsets[-1], 2)).astype(np.float16)
    x = rng.random((512, 3)).astype(np.float32)
    h, _ = oracle.grid_encode_forward(x, emb, offsets, pls, 16)
    f, _ = oracle.grid_encode_forward(x, emb.astype(np.float32), offsets, pls, 16)
    assert h.dtype == np.float16
    err = np.abs(h.astype(np.float32) - f)
    assert err.max() < 4e-3 and err.max() > 0      # differs from fp32 only by fp16 rounding


def test_grid_backward_is_adjoint_of_forward():
    # <grad, forward(emb)> == <backward(grad), emb>  (the encoder is linear in the table)
    rng = np.random.default_rng(11)
    pls = oracle.per_level_scale_for(512)
    offsets = oracle.grid_offsets(3, 8, pls, 16, 15)
    emb = rng.normal(size=(offsets[-1], 2)).astype(np.float32)
    x = rng.random((300, 3)).astype(np.float32)
    grad = rng.normal(size=(8, 300, 2)).astype(np.float32)
    out, dy_dx = oracle.grid_encode_forward(x, emb, offsets, pls, 16, calc_grad_inputs=True)
    gg = oracle.grid_encode_backward(grad, x, offsets, offsets[-1], 2, pls, 16)
    lhs = float((grad.astype(np.float64) * out).sum())
    rhs = float((gg * emb).sum())
    assert lhs == pytest.approx(rhs, rel=1e-4)
    # dy_dx against central differences of the oracle itself (interior of cells, coarse levels)
    eps = 1e-4
    gi = oracle.grid_input_backward(grad, dy_dx)
    for d in range(3):
        xp, xm = x.copy(), x.copy()
        xp[:, d] += eps
        xm[:, d] -= eps
        fp, _ = oracle.grid_encode_forward(xp.astype(np.float64).astype(np.float32), emb, offsets, pls, 16)
        fm, _ = oracle.grid_encode_forward(xm, emb, offsets, pls, 16)
        fd = ((fp.astype(np.float64) - fm) / (2 * eps) * grad).sum(axis=(0, 2))
        ok = np.abs(fd - gi[:, d]) < 5e-2 * (1 + np.abs(fd))
        assert ok.mean() > 0.9      # samples whose +-eps stencil crosses a cell face on a fine level disagree by construction


# ------------------------------------------------------------------------------ SH
def test_sh_closed_forms_match_scipy_and_are_orthonormal():
    rng = np.random.default_rng(3)
    d = rng.normal(size=(512, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    got = oracle.sh_encode(d.astype(np.float32), 4)
    want = oracle.sh_encode_scipy(d, 4)
    assert got[0, 0] == pytest.approx(0.28209479177387814)
    assert np.abs(got - want).max() < 2e-6
    # orthonormality on a Gauss-Legendre x uniform-phi product quadrature
    ct, wt = np.polynomial.legendre.leggauss(16)
    phi = (np.arange(32) + 0.5) * 2 * np.pi / 32
    st = np.sqrt(1 - ct ** 2)
    pts = np.stack([np.outer(st, np.cos(phi)), np.outer(st, np.sin(phi)), np.outer(ct, np.ones_like(phi))], -1).reshape(-1, 3)
    w = np.outer(wt, np.full(32, 2 * np.pi / 32)).reshape(-1)
    Y = oracle.sh_encode(pts.astype(np.float32), 4).astype(np.float64)
    gram = (Y * w[:, None]).T @ Y
    assert np.abs(gram - np.eye(16)).max() < 1e-5


# ------------------------------------------------------------------------------ compositing
def _torch_composite(sigmas, rgbs, deltas, counts):
    """float64 torch restatement with cumprod (the formula of nerf/renderer.py:232-234)"""
    ws, depth, image = [], [], []
    o = 0
    for c in counts:
        s, r, dl = sigmas[o:o + c], rgbs[o:o + c], deltas[o:o + c]
        alpha = 1 - torch.exp(-s * dl[:, 0])
        T = torch.cumprod(torch.cat([torch.ones(1, dtype=s.dtype), 1 - alpha]), 0)[:-1]
        w = alpha * T
        t = torch.cumsum(dl[:, 1], 0)
        ws.append(w.sum())
        depth.append((w * t).sum())
        image.append((w[:, None] * r).sum(0))
        o += c
    return torch.stack(ws), torch.stack(depth), torch.stack(image)


@pytest.mark.parametrize("n_ch", [1, 3])
def test_composite_train_forward_backward(n_ch):
    rng = np.random.default_rng(n_ch)
    counts = [0, 1, 5, 37, 64, 100, 0, 33]
    M = sum(counts) + 7
    rays = np.zeros((len(counts), 3), np.int32)
    o = 0
    for i, c in enumerate(counts):
        rays[i] = (len(counts) - 1 - i, o, c)     # ray ids permuted
        o += c
    sig = rng.uniform(0, 30, M).astype(np.float32)
    rgb = rng.random((M, n_ch)).astype(np.float32)
    dl = np.stack([np.full(M, 0.0034, np.float32), rng.uniform(0.003, 0.02, M).astype(np.float32)], -1)
    ws, depth, image = oracle.composite_rays_train_forward(sig, rgb, dl, rays)

    ts, tr, td = (torch.tensor(a, dtype=torch.float64, requires_grad=g) for a, g in ((sig, True), (rgb, True), (dl, False)))
    tws, tdepth, timage = _torch_composite(ts, tr, td, counts)
    ids = rays[:, 0]
    assert np.allclose(ws[ids], tws.detach().numpy(), atol=1e-5)
    assert np.allclose(depth[ids], tdepth.detach().numpy(), atol=1e-5)
    assert np.allclose(image[ids], timage.detach().numpy(), atol=1e-5)

    gws = rng.normal(size=len(counts)).astype(np.float32)
    gim = rng.normal(size=(len(counts), n_ch)).astype(np.float32)
    gs, gr = oracle.composite_rays_train_backward(gws, gim, sig, rgb, dl, rays, ws, image)
    loss = (tws * torch.tensor(gws[ids], dtype=torch.float64)).sum() + (timage * torch.tensor(gim[ids], dtype=torch.float64)).sum()
    loss.backward()
    assert np.allclose(gs, ts.grad.numpy(), atol=2e-5, rtol=1e-4)
    assert np.allclose(gr, tr.grad.numpy(), atol=1e-6, rtol=1e-5)


def test_composite_overflowing_ray_is_zeroed():
    # offset + num_steps >= M -> ray contributes nothing (raymarching.cu:521-528), note the `>=`
    rays = np.array([[0, 0, 4], [1, 4, 4]], np.int32)
    sig = np.ones(8, np.float32)
    rgb = np.ones((8, 3), np.float32)
    dl = np.full((8, 2), 0.1, np.float32)
    ws, _, image = oracle.composite_rays_train_forward(sig, rgb, dl, rays)
    assert ws[0] > 0 and ws[1] == 0 and np.all(image[1] == 0)


def test_composite_inference_and_compaction():
    rng = np.random.default_rng(9)
    N, n_alive, n_step = 16, 8, 4
    alive = rng.permutation(N)[:n_alive].astype(np.int32)
    t = rng.uniform(0.2, 1.0, n_alive).astype(np.float32)
    sig = rng.uniform(0, 50, n_alive * n_step).astype(np.float32)
    rgb = rng.random((n_alive * n_step, 3)).astype(np.float32)
    dl = np.full((n_alive * n_step, 2), 0.0034, np.float32)
    dl[5 * n_step + 2:6 * n_step] = 0                   # ray slot 5 ran out of samples after 2
    sig[3 * n_step:4 * n_step] = 1e4                    # ray slot 3 saturates -> terminated by T < 1e-5
    ws0, d0, im0 = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    t1, ws, d, im = oracle.composite_rays(n_alive, n_step, alive, t, sig, rgb, dl, ws0, d0, im0)
    assert t1[5] == -1 and t1[3] == -1
    live = [i for i in range(n_alive) if i not in (3, 5)]
    assert np.allclose(t1[live], t[live] + n_step * np.float32(0.0034), atol=1e-6)
    assert np.all(ws[alive] > 0) and np.all(ws[np.setdiff1d(np.arange(N), alive)] == 0)
    ra, rt, cnt = oracle.compact_rays(n_alive, alive, t1)
    assert cnt == n_alive - 2 and np.array_equal(ra[:cnt], alive[live]) and np.array_equal(rt[:cnt], t1[live])


# ------------------------------------------------------------------------------ marcher
@pytest.mark.parametrize("bound,perturb", [(1, False), (3, True)])
def test_marcher_invariants(bound, perturb):
    cascade = 1 + int(np.ceil(np.log2(bound)))
    grid = synthetic.ball_density_grid(bound, cascade)
    bits = synthetic.packbits_np(grid)
    o, d = synthetic.random_rays(64, bound, seed=1)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    xyzs, dirs, deltas, rays, counter = oracle.march_rays_train(o, d, bound, bits, cascade, 128, nears, fars, perturb=perturb)
    dt_min = np.float32(2 * 1.7320508075688772 / 1024)
    assert counter[1] == 64 and counter[0] == rays[:, 2].sum() and counter[0] > 64
    assert np.array_equal(rays[:, 0], np.arange(64))
    assert np.array_equal(rays[:, 1], np.concatenate([[0], np.cumsum(rays[:-1, 2])]))
    m = counter[0]
    assert np.all(deltas[:m, 0] == dt_min)
    assert np.all(deltas[:m, 1] >= dt_min * 0.999)
    assert np.all(np.abs(xyzs[:m]).max(-1) <= bound)
    assert np.all(xyzs[m:] == 0)
    # every emitted sample lies in a cell whose bit is set, and inside/near the ball
    r = np.linalg.norm(xyzs[:m], axis=-1)
    assert r.max() < 0.5 * bound + 2 * np.sqrt(3) * bound / 128 * 2
    for i in range(0, 64, 7):
        off, cnt = rays[i, 1], rays[i, 2]
        assert np.allclose(dirs[off:off + cnt], d[i])
        if cnt:
            t = np.linalg.norm(xyzs[off:off + cnt] - o[i], axis=-1)
            assert np.all(np.diff(t) > 0)
            first_t = t[0]
            assert first_t >= nears[i] - 1e-4 and t[-1] < fars[i]
            if perturb:
                jit = dt_min * np.float32(oracle.pcg32_first_float(i))
                k = np.round((first_t - nears[i] - jit) / dt_min)
                assert abs(nears[i] + jit + k * dt_min - first_t) < 2e-4     # on the t0 + k*dt lattice


def test_marcher_full_grid_emits_every_step_and_respects_max_steps():
    bits = np.full(128 ** 3 // 8, 255, np.uint8)
    o = np.array([[-0.9, 0.01, 0.02]], np.float32)
    d = np.array([[1.0, 0.0, 0.0]], np.float32) + 1e-9
    nears, fars = oracle.near_far_from_aabb(o, d, np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2)
    xyzs, _, deltas, rays, counter = oracle.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars)
    dt = np.float32(2 * 1.7320508075688772 / 1024)
    want = int(np.ceil((fars[0] - nears[0]) / dt))
    assert abs(int(rays[0, 2]) - want) <= 1
    assert np.allclose(np.diff(xyzs[:rays[0, 2], 0]), dt, atol=1e-6)
    # max_steps sets dt_min = 2*sqrt(3)/max_steps; clamp(x, dt_min, dt_max) = min(dt_max, max(dt_min, x)), so once
    # dt_min > dt_max (= 2*sqrt(3)/128 here) the step is dt_max (raymarching.cu:36-38,344-345,366)
    _, _, deltas2, rays2, _ = oracle.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars, max_steps=64)
    dt2 = np.float32(2 * np.float32(1.7320508075688772) / 128)
    assert deltas2[0, 0] == dt2
    assert abs(int(rays2[0, 2]) - min(64, int(np.ceil((fars[0] - nears[0]) / dt2)))) <= 1
    _, _, deltas3, rays3, _ = oracle.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars, max_steps=256)
    assert deltas3[0, 0] == np.float32(2 * np.float32(1.7320508075688772) / 256)
    assert rays3[0, 2] == min(256, int(np.ceil((fars[0] - nears[0]) / deltas3[0, 0])))


def test_marcher_inference_continues_where_training_marcher_goes():
    bound, cascade = 1, 1
    bits = synthetic.packbits_np(synthetic.ball_density_grid(bound, cascade))
    o, d = synthetic.random_rays(16, bound, seed=2)
    nears, fars = oracle.near_far_from_aabb(o, d, np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2)
    xyzs, _, deltas, rays, _ = oracle.march_rays_train(o, d, bound, bits, cascade, 128, nears, fars)
    alive = np.arange(16, dtype=np.int32)
    x2, _, dl2 = oracle.march_rays(16, 8, alive, nears, o, d, bound, bits, cascade, 128, nears, fars)
    for i in range(16):
        cnt = min(8, rays[i, 2])
        off = rays[i, 1]
        assert np.array_equal(x2[i * 8:i * 8 + cnt], xyzs[off:off + cnt])
        assert np.array_equal(dl2[i * 8:i * 8 + cnt], deltas[off:off + cnt])
        assert np.all(dl2[i * 8 + cnt:(i + 1) * 8] == 0)


# ------------------------------------------------------------------------------ FFMLP
@pytest.mark.parametrize("num_layers", [2, 3])
def test_ffmlp_oracle_matches_torch_autograd(num_layers):
    torch.manual_seed(0)
    B, I, W = 256, 32, 64
    n = W * (I + W * (num_layers - 1) + 16)
    w = (torch.rand(n) * 2 - 1) * (3 / W) ** 0.5
    x = torch.randn(B, I) * 0.5
    g = torch.randn(B, 16) * 0.1
    wh, xh, gh = w.half().numpy(), x.half().numpy(), g.half().numpy()
    y, fb = oracle.ffmlp_forward(xh, wh, I, W, num_layers)
    gx, gw, bb = oracle.ffmlp_backward(gh, xh, wh, fb, I, W, num_layers)

    mats = [torch.tensor(m.astype(np.float64), requires_grad=True) for m in oracle.ffmlp_split(wh, I, W, num_layers)]
    xt = torch.tensor(xh.astype(np.float64), requires_grad=True)
    h = xt
    for k in range(num_layers):
        h = torch.relu(h @ mats[k].T)
    yt = h @ mats[-1].T
    (yt * torch.tensor(gh.astype(np.float64))).sum().backward()
    assert np.abs(y - yt.detach().numpy()).max() < 2e-2          # fp16 storage of activations
    assert fb.shape == (num_layers, B, W) and bb.shape == (num_layers, B, W)
    gw_t = np.concatenate([m.grad.numpy().reshape(-1) for m in mats])
    assert np.abs(gw - gw_t).max() < 2e-2 * max(1.0, np.abs(gw_t).max())
    assert np.abs(gx - xt.grad.numpy()).max() < 2e-2

"""GPU parity: raymarching kernels (through the drop-in package -> C-ABI) vs the CPU oracle and,
when oracle/_ref is present, vs the reference's own CUDA build.  Integer/index outputs must be
bit-exact; the marcher's float outputs are bit-exact too (same fp32 operation sequence);
compositing sums are reassociated by the warp scans -> 1e-5 abs / 1e-4 rel (stated per test)."""
import numpy as np
import pytest
import torch

from enerf_b200 import raymarching as rm
from oracle import oracle
from tests.gpu_common import DEV, n, per_ray, ref_mod, scene, t

pytestmark = pytest.mark.gpu


def test_near_far_bit_exact():
    rng = np.random.default_rng(0)
    o = rng.uniform(-2, 2, (5000, 3)).astype(np.float32)
    d = rng.normal(size=(5000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d[:10, 0] = 0.0                                             # axis-parallel rays: 1/0 = inf
    aabb = np.array([-1, -1.5, -1, 1, 1.5, 1], np.float32)
    wn, wf = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    gn, gf = rm.near_far_from_aabb(t(o), t(d), t(aabb), 0.2)
    ok = ~(np.isnan(wn) | np.isnan(wf))
    assert np.array_equal(n(gn)[ok], wn[ok]) and np.array_equal(n(gf)[ok], wf[ok])
    assert (wn == np.float32(3.4028234663852886e38)).sum() > 100        # misses exercised
    R = ref_mod("_raymarching")
    if R is not None:
        rn, rf = torch.empty(5000, device=DEV), torch.empty(5000, device=DEV)
        R.near_far_from_aabb(t(o), t(d), t(aabb), 5000, 0.2, rn, rf)
        assert torch.equal(rn[torch.from_numpy(ok).to(DEV)], gn[torch.from_numpy(ok).to(DEV)])
        assert torch.equal(rf[torch.from_numpy(ok).to(DEV)], gf[torch.from_numpy(ok).to(DEV)])


def test_morton_packbits_polar():
    rng = np.random.default_rng(1)
    coords = rng.integers(0, 128, (10007, 3)).astype(np.int32)
    idx = rm.morton3D(t(coords))
    assert np.array_equal(n(idx), oracle.morton3D(coords))
    assert np.array_equal(n(rm.morton3D_invert(idx)), coords)
    grid = rng.random((3, 128 ** 3)).astype(np.float32) - 0.3
    grid[1, :4096] = -1.0
    assert np.array_equal(n(rm.packbits(t(grid), 0.2)), oracle.packbits(grid, 0.2))
    small = rng.random((1, 64)).astype(np.float32)               # 8 bytes: less than one warp of work
    assert np.array_equal(n(rm.packbits(t(small), 0.5)), oracle.packbits(small, 0.5))
    R = ref_mod("_raymarching")
    if R is not None:
        o = rng.uniform(-0.5, 0.5, (1000, 3)).astype(np.float32)
        d = rng.normal(size=(1000, 3)).astype(np.float32)
        want = torch.empty(1000, 2, device=DEV)
        R.polar_from_ray(t(o), t(d), 2.0, 1000, want)
        got = rm.polar_from_ray(t(o), t(d), 2.0)
        assert torch.allclose(got, want, atol=1e-5)


def _march_case(bound, perturb, dt_gamma, n_rays, max_steps=1024, seed=0):
    sc = scene(n_rays, bound, seed)
    want = oracle.march_rays_train(sc["o"], sc["d"], bound, sc["bits"], sc["cascade"], 128, sc["nears"], sc["fars"], perturb=perturb,
                                   dt_gamma=dt_gamma, max_steps=max_steps)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    got = rm.march_rays_train(t(sc["o"]), t(sc["d"]), float(bound), t(sc["bits"]), sc["cascade"], 128, t(sc["nears"]), t(sc["fars"]),
                              counter, -1, perturb, 128, False, dt_gamma, max_steps)
    return sc, want, got, counter


@pytest.mark.parametrize("bound,perturb,dt_gamma,n_rays", [(1, False, 0.0, 1000), (3, True, 0.0, 1021), (2, True, 1.0 / 128, 515),
                                                           (3, False, 1.0 / 256, 300)])
def test_march_rays_train_matches_oracle_per_ray(bound, perturb, dt_gamma, n_rays):
    sc, (wx, wd, wdl, wrays, wcnt), (gx, gd, gdl, grays), counter = _march_case(bound, perturb, dt_gamma, n_rays)
    gx, gd, gdl, grays = n(gx), n(gd), n(gdl), n(grays)
    assert n(counter).tolist() == wcnt.tolist()
    assert sorted(grays[:, 0].tolist()) == list(range(n_rays))                  # every ray has exactly one row
    m = int(wcnt[0])
    assert gx.shape[0] == m + (128 - m % 128) and np.all(gx[m:] == 0)           # padded like raymarching.py:220-224
    # ranges tile [0, m) without overlap
    order = np.argsort(grays[:, 1], kind="stable")
    offs, cnts = grays[order, 1], grays[order, 2]
    assert offs[0] == 0 and np.array_equal(offs[1:], np.cumsum(cnts)[:-1])
    W = per_ray(wrays, wx, wd, wdl)
    G = per_ray(grays, gx, gd, gdl)
    bad = [r for r in range(n_rays) if len(W[r][0]) != len(G[r][0])]
    assert not bad, f"{len(bad)} rays with a different sample count, e.g. ray {bad[0]}: oracle {len(W[bad[0]][0])} vs gpu {len(G[bad[0]][0])}"
    for r in range(n_rays):
        for a, b, name in zip(W[r], G[r], ("xyzs", "dirs", "deltas")):
            assert np.array_equal(a, b), f"ray {r} {name}: max diff {np.abs(a - b).max()}"


def test_march_rays_train_matches_reference_build():
    R = ref_mod("_raymarching")
    if R is None:
        pytest.skip("oracle/_ref not built")
    for bound, perturb, dt_gamma in [(3, True, 0.0), (1, False, 0.0), (2, True, 1.0 / 128)]:
        n_rays = 2048
        sc = scene(n_rays, bound, seed=4)
        o, d, bits, nears, fars = t(sc["o"]), t(sc["d"]), t(sc["bits"]), t(sc["nears"]), t(sc["fars"])
        M = n_rays * 1024
        rx, rd, rdl = torch.zeros(M, 3, device=DEV), torch.zeros(M, 3, device=DEV), torch.zeros(M, 2, device=DEV)
        rrays = torch.empty(n_rays, 3, dtype=torch.int32, device=DEV)
        rc = torch.zeros(2, dtype=torch.int32, device=DEV)
        R.march_rays_train(o, d, bits, float(bound), dt_gamma, 1024, n_rays, sc["cascade"], 128, M, nears, fars, rx, rd, rdl, rrays, rc, int(perturb))
        gc = torch.zeros(2, dtype=torch.int32, device=DEV)
        gx, gd, gdl, grays = rm.march_rays_train(o, d, float(bound), bits, sc["cascade"], 128, nears, fars, gc, -1, perturb, 128, False, dt_gamma, 1024)
        assert torch.equal(rc, gc)
        W = per_ray(n(rrays), n(rx), n(rd), n(rdl))
        G = per_ray(n(grays), n(gx), n(gd), n(gdl))
        n_bad = 0
        for r in range(n_rays):
            same = len(W[r][0]) == len(G[r][0]) and all(np.array_equal(a, b) for a, b in zip(W[r], G[r]))
            n_bad += not same
        assert n_bad == 0, f"bound {bound}: {n_bad}/{n_rays} rays differ from the reference build"


def test_march_rays_train_mean_count_overflow_drops_rays():
    sc = scene(512, 1, seed=2)
    o, d, bits, nears, fars = t(sc["o"]), t(sc["d"]), t(sc["bits"]), t(sc["nears"]), t(sc["fars"])
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    full = rm.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars, counter, -1, False, 128, False, 0, 1024)
    total = int(counter[0])
    mean_count = total // 2
    counter.zero_()
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars, counter, mean_count, False, 128, False, 0, 1024)
    M = mean_count + (128 - mean_count % 128)
    assert xyzs.shape[0] == M and int(counter[0]) == total            # counter still counts everything
    rays = n(rays)
    kept = rays[rays[:, 1] + rays[:, 2] < M]
    dropped = rays[rays[:, 1] + rays[:, 2] >= M]
    assert len(dropped) > 0 and len(kept) > 0
    F = per_ray(n(full[3]), n(full[0]))
    xs = n(xyzs)
    for rid, off, cnt in kept[:50]:
        assert np.array_equal(xs[off:off + cnt], F[int(rid)][0])
    # composite zeroes the dropped rays
    sig = torch.ones(M, device=DEV)
    rgb = torch.ones(M, 3, device=DEV)
    ws, depth, image = rm.composite_rays_train(sig, rgb, deltas, t(rays))
    assert np.all(n(ws)[dropped[:, 0]] == 0) and np.all(n(image)[dropped[:, 0]] == 0)


@pytest.mark.parametrize("n_ch", [1, 3])
def test_composite_train_forward_backward(n_ch):
    sc, (wx, wd, wdl, wrays, wcnt), _, _ = _march_case(3, True, 0.0, 700, seed=3)
    rng = np.random.default_rng(7)
    M = int(wcnt[0]) + 5
    sig = (rng.uniform(0, 40, M) * (rng.random(M) < 0.7)).astype(np.float32)
    rgb = rng.random((M, n_ch)).astype(np.float32)
    dl = wdl[:M]
    ws, depth, image = oracle.composite_rays_train_forward(sig, rgb, dl, wrays)
    ts, tr = t(sig).requires_grad_(True), t(rgb).requires_grad_(True)
    gws, gdepth, gimage = rm.composite_rays_train(ts, tr, t(dl), t(wrays))
    for a, b, name in ((gws, ws, "weights_sum"), (gdepth, depth, "depth"), (gimage, image, "image")):
        assert np.allclose(n(a), b, atol=1e-5, rtol=1e-4), f"{name}: {np.abs(n(a) - b).max()}"
    g1 = rng.normal(size=ws.shape).astype(np.float32)
    g3 = rng.normal(size=image.shape).astype(np.float32)
    (gws * t(g1)).sum().backward(retain_graph=True)
    (gimage * t(g3)).sum().backward()
    wgs, wgr = oracle.composite_rays_train_backward(g1, g3, sig, rgb, dl, wrays, ws, image)
    assert np.allclose(n(ts.grad), wgs, atol=2e-5, rtol=1e-3), np.abs(n(ts.grad) - wgs).max()
    assert np.allclose(n(tr.grad), wgr, atol=1e-6, rtol=1e-4)
    R = ref_mod("_raymarching")
    if R is not None and n_ch == 3:
        N = wrays.shape[0]
        rws, rdp, rim = torch.empty(N, device=DEV), torch.empty(N, device=DEV), torch.empty(N, 3, device=DEV)
        R.composite_rays_train_forward(t(sig), t(rgb), t(dl), t(wrays), M, N, rws, rdp, rim)
        assert torch.allclose(rws, gws.detach(), atol=1e-5, rtol=1e-4) and torch.allclose(rim, gimage.detach(), atol=1e-5, rtol=1e-4)
        assert torch.allclose(rdp, gdepth.detach(), atol=1e-5, rtol=1e-4)
        rgs, rgr = torch.zeros(M, device=DEV), torch.zeros(M, 3, device=DEV)
        R.composite_rays_train_backward(t(g1), t(g3), t(sig), t(rgb), t(dl), t(wrays), rws, rim, M, N, rgs, rgr)
        assert torch.allclose(rgs, ts.grad, atol=2e-5, rtol=1e-3) and torch.allclose(rgr, tr.grad, atol=1e-6, rtol=1e-4)


def test_inference_loop_primitives_match_oracle():
    bound = 2
    sc = scene(999, bound, seed=5)
    N = 999
    o, d, bits, nears, fars = t(sc["o"]), t(sc["d"]), t(sc["bits"]), t(sc["nears"]), t(sc["fars"])
    rng = np.random.default_rng(3)
    alive_np = rng.permutation(N)[:700].astype(np.int32)
    t_np = sc["nears"][alive_np].copy()
    for n_step, perturb in [(1, 0), (4, 0), (8, 3)]:
        wx, wd, wdl = oracle.march_rays(700, n_step, alive_np, t_np, sc["o"], sc["d"], bound, sc["bits"], sc["cascade"], 128, sc["nears"],
                                        sc["fars"], perturb=perturb)
        gx, gd, gdl = rm.march_rays(700, n_step, t(alive_np), t(t_np), o, d, float(bound), bits, sc["cascade"], 128, nears, fars, 128, perturb, 0, 1024)
        m = 700 * n_step
        assert gx.shape[0] == m + (128 - m % 128)
        assert np.array_equal(n(gx)[:m], wx) and np.array_equal(n(gd)[:m], wd) and np.array_equal(n(gdl)[:m], wdl)
        assert np.all(n(gx)[m:] == 0)
        sig = (rng.uniform(0, 200, m)).astype(np.float32)
        rgb = rng.random((m, 3)).astype(np.float32)
        ws0, d0, im0 = rng.random(N).astype(np.float32) * 0.5, rng.random(N).astype(np.float32), rng.random((N, 3)).astype(np.float32)
        wt, wws, wdp, wim = oracle.composite_rays(700, n_step, alive_np, t_np, sig, rgb, wdl, ws0, d0, im0)
        gt, gws_, gdp, gim = t(t_np), t(ws0), t(d0), t(im0)
        rm.composite_rays(700, n_step, t(alive_np), gt, t(sig), t(rgb), gdl, gws_, gdp, gim)
        assert np.array_equal(n(gt) < 0, wt < 0)
        assert np.allclose(n(gt), wt, atol=1e-6) and np.allclose(n(gws_), wws, atol=1e-6) and np.allclose(n(gdp), wdp, atol=1e-5)
        assert np.allclose(n(gim), wim, atol=1e-6)
        # compaction: same survivor set (slot order is arbitrary in the reference; ours is stable per warp)
        ra, rt = torch.zeros(N, dtype=torch.int32, device=DEV), torch.zeros(N, device=DEV)
        cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        rm.compact_rays(700, ra, t(alive_np), rt, gt, cnt)
        wra, wrt, wc = oracle.compact_rays(700, alive_np, wt)
        assert int(cnt[0]) == wc
        got = dict(zip(n(ra)[:wc].tolist(), n(rt)[:wc].tolist()))
        want = dict(zip(wra[:wc].tolist(), wrt[:wc].tolist()))
        assert got.keys() == want.keys()
        assert all(abs(got[k] - want[k]) < 1e-6 for k in want)


@pytest.mark.parametrize("live", [700, 123, 0])
def test_inference_march_writes_every_row_of_its_slots(live):
    """k_march_rays leaves nothing of the n_alive * n_step rows to the caller: samples, zeros behind a ray's last sample (the padding
    composite_rays stops at), zeros in the slots at or beyond the device-side alive count — on buffers that start out as NaN"""
    from enerf_b200.backends import raymarching_backend as RB
    bound, N, n_alive, n_step = 2, 999, 700, 26
    sc = scene(N, bound, seed=5)
    o, d, bits, nears, fars = t(sc["o"]), t(sc["d"]), t(sc["bits"]), t(sc["nears"]), t(sc["fars"])
    alive_np = np.random.default_rng(3).permutation(N)[:n_alive].astype(np.int32)
    t_np = sc["nears"][alive_np].copy()
    t_np[::7] = sc["fars"][alive_np][::7] - np.float32(0.05)      # rays about to end: fewer than n_step samples, then padding
    wx, wd, wdl = oracle.march_rays(live, n_step, alive_np, t_np, sc["o"], sc["d"], bound, sc["bits"], sc["cascade"], 128, sc["nears"], sc["fars"])
    M = n_alive * n_step
    xyzs, dirs, deltas = (torch.full((M, k), float("nan"), device=DEV) for k in (3, 3, 2))
    count = torch.tensor([live], dtype=torch.int32, device=DEV)
    for ob in (None, RB.occupancy_bounds(bits, sc["cascade"], 128)):
        for buf in (xyzs, dirs, deltas):
            buf.fill_(float("nan"))
        RB.march_rays(n_alive, n_step, t(alive_np), t(t_np), o, d, float(bound), 0.0, 1024, sc["cascade"], 128, bits, nears, fars, xyzs, dirs, deltas, 0,
                      count, ob)
        m = live * n_step
        assert np.array_equal(n(xyzs)[:m], wx) and np.array_equal(n(dirs)[:m], wd) and np.array_equal(n(deltas)[:m], wdl)
        assert 0 < np.count_nonzero(wdl[:, 0] == 0) < max(m, 1) or live == 0      # the case has padding rows
        assert bool((xyzs[m:] == 0).all()) and bool((dirs[m:] == 0).all()) and bool((deltas[m:] == 0).all())


def test_full_size_properties_4096_rays_bound3():
    """BASELINE config-2 shape: size-independent invariants instead of an oracle pass."""
    sc = scene(4096, 3, seed=11)
    o, d, bits, nears, fars = t(sc["o"]), t(sc["d"]), t(sc["bits"]), t(sc["nears"]), t(sc["fars"])
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, 3.0, bits, 3, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
    m = int(counter[0])
    assert int(counter[1]) == 4096 and int(rays[:, 2].sum()) == m
    dt_min = np.float32(2 * np.float32(1.7320508075688772) / 1024)
    assert torch.all(deltas[:m, 0] == float(dt_min)) and torch.all(deltas[:m, 1] >= float(dt_min) * 0.999)
    r = xyzs[:m].norm(dim=-1)
    assert float(r.max()) < 1.5 + 0.25          # inside the ball (+ one coarse cell)
    # idempotence: marching again gives the same per-ray counts
    counter.zero_()
    _, _, _, rays2 = rm.march_rays_train(o, d, 3.0, bits, 3, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
    c1 = torch.zeros(4096, dtype=torch.int64, device=DEV).scatter_(0, rays[:, 0].long(), rays[:, 2].long())
    c2 = torch.zeros(4096, dtype=torch.int64, device=DEV).scatter_(0, rays2[:, 0].long(), rays2[:, 2].long())
    assert torch.equal(c1, c2)
    # compositing a constant colour gives image == weights_sum * colour and ws in [0,1]
    sig = torch.rand(xyzs.shape[0], device=DEV) * 20
    rgb = torch.full((xyzs.shape[0], 3), 0.25, device=DEV)
    ws, depth, image = rm.composite_rays_train(sig, rgb, deltas, rays)
    assert torch.allclose(image, ws[:, None] * 0.25, atol=1e-5)
    assert float(ws.min()) >= 0 and float(ws.max()) <= 1 + 1e-5


# ---- marching bounded by the occupied box (enerf_occupancy_bounds + *_bounded): same samples as the exhaustive march ----------
def _bits_from_cells(C, H, cells):
    """bitfield with exactly the given (level, x, y, z) cells set"""
    grid = np.zeros((C, H ** 3), np.float32)
    for lv, x, y, z in cells:
        grid[lv, int(oracle.morton3D(np.array([[x, y, z]], np.int32))[0])] = 1.0
    return oracle.packbits(grid, 0.5), grid


def _cell_box(bits, C, H):
    """numpy restatement of the bounds kernel: 128 consecutive Morton codes (an 8 x 4 x 4 block) count as a whole"""
    out = np.zeros((C, 6), np.int32)
    cells = np.unpackbits(np.asarray(bits, np.uint8), bitorder="little").reshape(C, -1)
    for lv in range(C):
        occ = np.nonzero(cells[lv].reshape(-1, 128).any(axis=1))[0]
        if occ.size == 0:
            out[lv] = [H, H, H, -1, -1, -1]
            continue
        xyz = oracle.morton3D_invert((occ * 128).astype(np.int32))
        out[lv, :3] = xyz.min(axis=0)
        out[lv, 3:] = np.minimum((xyz + np.array([7, 3, 3])).max(axis=0), H - 1)
    return out


@pytest.mark.parametrize("case", ["ball", "sparse", "corner", "empty", "full", "dt_gamma"])
def test_bounded_march_emits_the_samples_of_the_exhaustive_march(case):
    from enerf_b200.backends import raymarching_backend as RB
    bound, C, H, N = 2, 2, 128, 3000
    rng = np.random.default_rng(5)
    if case in ("ball", "dt_gamma"):
        sc = scene(N, bound, seed=4)
        bits = sc["bits"]
    else:
        cells = {"sparse": [(int(rng.integers(0, C)), *rng.integers(20, 108, 3)) for _ in range(40)],
                 "corner": [(1, 0, 0, 0), (1, 127, 127, 127), (0, 0, 127, 64)], "empty": [],
                 "full": None}[case]
        if cells is None:
            grid = np.ones((C, H ** 3), np.float32)
            bits = oracle.packbits(grid, 0.5)
        else:
            bits, _ = _bits_from_cells(C, H, cells)
    from enerf_b200 import synthetic
    o, d = synthetic.random_rays(N, bound, seed=9)
    d[:8, 1] = 0.0                                                # axis-parallel rays
    d[:8] /= np.linalg.norm(d[:8], axis=-1, keepdims=True)
    o[8:11] = [[0.1, 0.1, 0.1], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]]  # three rays through the corner cells of the "corner" case
    d[8:11] = [[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [-1.0, 1.0, 0.008]]
    d[8:11] /= np.linalg.norm(d[8:11], axis=-1, keepdims=True)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    tb = t(bits)
    bounds = RB.occupancy_bounds(tb, C, H)
    box = _cell_box(bits, C, H)
    assert np.array_equal(n(bounds), box)
    # the float rows behind the integer rows (ENERF_OCC_BOUNDS_WORDS, include/enerf_b200.h): the same box in units of the level's half extent
    off, words = (6 * C + 3) & ~3, ((6 * C + 3) & ~3) + 8 * C
    flat = torch.empty(0, dtype=torch.int32, device=DEV).set_(bounds.untyped_storage(), 0, (words,))
    rows = n(flat[off:].view(torch.float32)).reshape(C, 8)
    want = np.zeros((C, 8), np.float32)
    cell = np.float32(2.0) / np.float32(H)
    want[:, 0:3] = box[:, :3].astype(np.float32) * cell - np.float32(1.0)       # exact: H is a power of two
    want[:, 4:7] = (box[:, 3:] + 1).astype(np.float32) * cell - np.float32(1.0)
    want[:, 3] = (box[:, 3] >= box[:, 0]).astype(np.float32)
    assert np.array_equal(rows, want)
    dt_gamma = 1.0 / 128 if case == "dt_gamma" else 0.0
    res = []
    M = N * 1024
    for ob in (None, bounds):
        xyzs, dirs, deltas = (torch.zeros(M, k, device=DEV) for k in (3, 3, 2))
        rays = torch.empty(N, 3, dtype=torch.int32, device=DEV)
        counter = torch.zeros(2, dtype=torch.int32, device=DEV)
        RB.march_rays_train(t(o), t(d), tb, float(bound), dt_gamma, 1024, N, C, H, M, t(nears), t(fars), xyzs, dirs, deltas, rays, counter, 1, ob)
        res.append((int(counter[0]), per_ray(n(rays), n(xyzs), n(deltas))))
    assert res[0][0] == res[1][0]
    if case == "empty":
        assert res[0][0] == 0
    elif case == "corner":
        # (the ray into the +++ corner emits nothing: the reference's skip target uses (H-1) as divisor, raymarching.cu:391-393, and jumps
        # from cell 126 past cell 127 — reproduced, see oracle.march_rays_train)
        assert res[0][0] > 0 and len(res[0][1][9][0]) > 0 and len(res[0][1][10][0]) > 0
    elif case != "sparse":
        assert res[0][0] > 1000
    for rid, (x0, dl0) in res[0][1].items():
        x1, dl1 = res[1][1][rid]
        assert np.array_equal(x0, x1) and np.array_equal(dl0, dl1), rid
    # inference rounds: 3 rounds of 40 steps from the same state, with and without bounds
    outs = []
    for ob in (None, bounds):
        rays_alive = torch.arange(N, dtype=torch.int32, device=DEV)
        rays_t = t(nears).clone()
        acc = []
        for _ in range(3):
            xyzs, dirs, deltas = rm.march_rays(N, 40, rays_alive, rays_t, t(o), t(d), float(bound), tb, C, H, t(nears), t(fars), 128, False, dt_gamma, 1024,
                                               None, ob)
            acc.append((xyzs.clone(), deltas.clone()))
            # advance every ray by what it consumed, like composite_rays does (sum of deltas[:, 1]); rays that ended keep their t
            adv = deltas[:N * 40, 1].view(N, 40).sum(dim=1)
            rays_t = rays_t + adv
        outs.append(acc)
    for (xa, da), (xb, db) in zip(*outs):
        assert torch.equal(xa, xb) and torch.equal(da, db)


@pytest.mark.parametrize("n_ch,bg_kind", [(1, "scalar"), (3, "scalar"), (3, "vector"), (3, "per_ray"), (4, "zero_dim"), (1, "one_by_one")])
def test_finish_rays_is_the_aten_expression(n_ch, bg_kind):
    """renderer.py:397-398 as one kernel: forward bit-identical to the ATen expression, gradients to 1e-6"""
    g = torch.Generator(device=DEV).manual_seed(3)
    N = 5000
    ws = torch.rand(N, device=DEV, generator=g).requires_grad_()
    depth = (torch.rand(N, device=DEV, generator=g) * 3).requires_grad_()
    image = torch.rand(N, n_ch, device=DEV, generator=g).requires_grad_()
    nears = torch.rand(N, device=DEV, generator=g) + 0.2
    fars = nears + torch.rand(N, device=DEV, generator=g) * 4 + 0.1
    nears[:5] = fars[:5] = 3.4028234663852886e38                       # rays that miss the box: 0 / 0
    bg = {"scalar": 1, "vector": torch.rand(n_ch, device=DEV, generator=g), "per_ray": torch.rand(N, n_ch, device=DEV, generator=g),
          "zero_dim": torch.tensor(0.25, device=DEV), "one_by_one": torch.rand(1, 1, 1, device=DEV, generator=g)}[bg_kind]
    gi, gd = torch.randn(N, n_ch, device=DEV, generator=g), torch.randn(N, device=DEV, generator=g)
    gd[:5] = 0
    outs = []
    for fused in (True, False):
        for v in (ws, depth, image):
            v.grad = None
        if fused:
            im, dp = rm.finish_rays(ws, depth, image, nears, fars, bg)
        else:
            im, dp = image + (1 - ws).unsqueeze(-1) * bg, torch.clamp(depth - nears, min=0) / (fars - nears)
        im = im.reshape(N, n_ch)                                       # a [1, 1, 1] background broadcasts the ATen result to [1, N, n_ch]
        ((im * gi).sum() + (dp[5:] * gd[5:]).sum()).backward()
        outs.append((im.detach(), dp.detach(), ws.grad.clone(), depth.grad.clone(), image.grad.clone()))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1][5:], b[1][5:]) and bool(torch.isnan(a[1][:5]).all()) and bool(torch.isnan(b[1][:5]).all())
    assert float((a[2] - b[2]).abs().max()) <= 1e-5 and torch.equal(a[4], b[4])
    assert float((a[3][5:] - b[3][5:]).abs().max()) <= 1e-6 * float(b[3][5:].abs().max())


@pytest.mark.parametrize("seed", range(8))
def test_bounded_march_on_random_grids(seed):
    """random occupancy (salt of density 1e-5 .. 0.5, or one compact blob), bounds 1 / 2 / 3 / 0.75, rays that start anywhere in or around
    the volume, some axis-parallel, jitter on and off, constant and growing steps: per ray the same sample count, positions and deltas"""
    from enerf_b200.backends import raymarching_backend as RB
    H, N = 128, 3000
    rng = np.random.default_rng(seed)
    bound = [1, 2, 3, 0.75][seed % 4]
    C = 1 + int(np.ceil(np.log2(bound))) if bound > 1 else 1
    grid = (rng.random((C, H ** 3)) < 10 ** rng.uniform(-5, -0.3)).astype(np.float32)
    if seed % 5 == 0:
        grid[:] = 0
        c0 = rng.integers(10, 100, 3)
        idx = np.stack(np.meshgrid(*[np.arange(c0[k], c0[k] + rng.integers(1, 20)) for k in range(3)], indexing="ij"), -1).reshape(-1, 3).astype(np.int32)
        grid[rng.integers(0, C), oracle.morton3D(idx)] = 1
    bits = t(oracle.packbits(grid, 0.5))
    o = rng.uniform(-bound * 1.2, bound * 1.2, (N, 3)).astype(np.float32)
    d = rng.normal(size=(N, 3)).astype(np.float32)
    d[:20, rng.integers(0, 3)] = 0
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o, d = t(o), t(d)
    nears, fars = rm.near_far_from_aabb(o, d, t(np.array([-bound] * 3 + [bound] * 3, np.float32)), 0.05)
    dt_gamma, perturb = (0.0 if seed % 3 else 1.0 / 256), seed % 2
    bounds = RB.occupancy_bounds(bits, C, H)
    M, outs = N * 1024, []
    for ob in (None, bounds):
        xyzs, dirs, deltas = (torch.zeros(M, k, device=DEV) for k in (3, 3, 2))
        rays = torch.empty(N, 3, dtype=torch.int32, device=DEV)
        counter = torch.zeros(2, dtype=torch.int32, device=DEV)
        RB.march_rays_train(o, d, bits, float(bound), dt_gamma, 1024, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, perturb, ob)
        r = rays[torch.argsort(rays[:, 0])]
        total = int(counter[0])
        starts = torch.cumsum(r[:, 2], 0) - r[:, 2]
        src = torch.repeat_interleave(r[:, 1].long() - starts.long(), r[:, 2].long()) + torch.arange(total, device=DEV)      # ray-ordered rows
        outs.append((r[:, 2].clone(), xyzs[src].clone(), deltas[src].clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])

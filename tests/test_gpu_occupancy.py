"""Occupancy-grid maintenance against golden vectors produced by the reference's OWN Python
(`NeRFRenderer.mark_untrained_grid` / `update_extra_state`, nerf/renderer.py:408-563, executed on the CPU with a scripted density and
scripted RNG draws: tests/golden/make_golden_grid_state.py).  `density_grid` must be BIT-identical (64-bit checksum over all
C*128^3 cells) after every refresh, full and partial; the bitfield may differ in a handful of cells that sit within rounding of the
mean-density threshold (the reference sums the mean in fp32, the kernel in fp64); `mark_untrained_grid` may differ in cells exactly on
a frustum boundary (the reference's batched matmul rounds differently)."""
import os

import numpy as np
import pytest
import torch

from enerf_b200 import raymarching as rm
from enerf_b200.nerf.renderer import NeRFRenderer
from oracle import grid_state as gs
from oracle import oracle
from tests.gpu_common import DEV, n, t

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "grid_state.npz")
H = 128


def _renderer(bound):
    r = NeRFRenderer(bound=bound, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1).to(DEV)
    r._grid_density = lambda x: gs.scripted_density(x)
    return r


def _check(r, gold, tag, max_bit_flips=16):
    g = n(r.density_grid)
    assert gs.checksum64(g) == int(gold[f"{tag}_checksum"]), (tag, "density_grid is not bit-identical to the reference's",
                                                              float(np.abs(g.reshape(-1)[::127] - gold[f"{tag}_sample"]).max()))
    assert np.array_equal(g.reshape(-1)[::127], gold[f"{tag}_sample"])
    assert int((g < 0).sum()) == int(gold[f"{tag}_n_neg"]) and int((g > 0).sum()) == int(gold[f"{tag}_n_pos"])
    mean = r.mean_density
    assert abs(mean - float(gold[f"{tag}_mean_density"])) <= 1e-6 * float(gold[f"{tag}_mean_density"]), (tag, mean)
    bits = n(r.density_bitfield)
    # exact w.r.t. our own threshold ...
    assert np.array_equal(bits, oracle.packbits(g.reshape(-1), min(np.float32(mean), np.float32(0.01))))
    # ... and within a few borderline cells of the reference's
    flips = int(np.unpackbits(bits ^ gold[f"{tag}_bitfield"]).sum())
    assert flips <= max_bit_flips, (tag, flips)


def test_mark_untrained_and_refreshes_reproduce_the_reference():
    gold = np.load(GOLD)
    bound = int(gold["bound"])
    r = _renderer(bound)
    C, cells = r.cascade, H ** 3
    # ---- mark_untrained_grid
    r.mark_untrained_grid(torch.from_numpy(gold["poses"]), gold["intrinsic"])
    neg = (n(r.density_grid).reshape(-1) < 0)
    want = np.unpackbits(gold["untrained_bits"], bitorder="little").astype(bool)[:neg.size]
    assert int((neg != want).sum()) <= 64, int((neg != want).sum())
    assert 0.05 < neg.mean() < 0.5
    # continue from the reference's exact mask so that the refreshes can be compared bit for bit
    r.density_grid.copy_(t(np.where(want, -1.0, 0.0).astype(np.float32)).view(C, cells))
    r.step_counter[:3, 0] = torch.tensor([1000, 1300, 1100], dtype=torch.int32, device=DEV)
    r.local_step = 3
    # ---- two full refreshes (iter_density < 16)
    m = np.arange(cells, dtype=np.int64)
    for upd in range(2):
        noise = np.concatenate([gs.full_noise(upd, cas, m) for cas in range(C)], axis=0)
        r.update_extra_state(_draws={"noise": torch.from_numpy(noise)})
        _check(r, gold, f"full{upd}")
        if upd == 0:
            assert r.mean_count == int(gold["mean_count"]) and r.local_step == 0
    assert r.iter_density == 2
    # ---- two partial refreshes
    r.iter_density = 16
    for upd in range(2):
        rc, ro, nz = [], [], []
        for cas in range(C):
            n_occ = int((r.density_grid[cas] > 0).sum())
            c, o, u = gs.partial_draws(upd, cas, cells // 4, H, n_occ)
            rc.append(c), ro.append(o), nz.append(u)
        draws = {"rand_coords": torch.from_numpy(np.stack(rc)), "rand_occ": torch.from_numpy(np.stack(ro)),
                 "noise": torch.from_numpy(np.concatenate(nz, axis=0))}
        r.update_extra_state(_draws=draws)
        _check(r, gold, f"partial{upd}")


def test_refresh_with_the_in_kernel_rng_is_consistent():
    """production path (no scripted draws): the bitfield is packbits(grid, min(mean, thresh)), untrained cells stay -1, the same
    torch seed gives the same grid and another seed another one; jittered positions stay inside their cells."""
    base = _renderer(2)
    C, cells = base.cascade, H ** 3
    base.density_grid[0, ::7] = -1.0
    grids = []
    for seed in (1, 1, 2):
        r2 = _renderer(2)
        r2.density_grid.copy_(base.density_grid)
        torch.manual_seed(seed)
        r2.update_extra_state()
        for _ in range(2):
            r2.iter_density = 16
            r2.update_extra_state()
        g = n(r2.density_grid)
        assert (g[0, ::7] == -1.0).all()
        mean = r2.mean_density
        assert np.array_equal(n(r2.density_bitfield), oracle.packbits(g.reshape(-1), min(np.float32(mean), np.float32(0.01))))
        assert abs(mean - np.clip(g, 0, None).astype(np.float64).mean()) < 1e-7
        grids.append(g)
    assert np.array_equal(grids[0], grids[1]) and not np.array_equal(grids[0], grids[2])
    from enerf_b200 import _lib
    xyz = torch.empty(C * cells, 3, device=DEV)
    _lib.call("enerf_occ_points_full", _lib.ptr(xyz), C, H, 2.0, None, 1234, _lib.stream())
    coords = rm.morton3D_invert(torch.arange(cells, dtype=torch.int32, device=DEV)).float()
    for cas in range(C):
        b = min(2 ** cas, 2)
        centre = (2 * coords / (H - 1) - 1) * (b - b / H)
        assert float((xyz[cas * cells:(cas + 1) * cells] - centre).abs().max()) <= b / H * (1 + 1e-5)


@pytest.mark.parametrize("size", [0, 1, 15, 16, 4095, 4096, 4097, 1 << 20, (1 << 21) + 77])
def test_ordered_compaction_matches_nonzero(size):
    from enerf_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(size)
    v = torch.rand(size, device=DEV, generator=g)
    for thresh in (0.5, -1.0, 2.0):
        idx = torch.full((max(size, 1),), -7, dtype=torch.int32, device=DEV)
        cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        scratch = torch.empty(size // 4096 + 2, dtype=torch.int32, device=DEV)
        _lib.call("enerf_compact_greater", _lib.ptr(v), thresh, size, _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(scratch), _lib.stream())
        want = torch.nonzero(v > thresh).squeeze(-1).int()
        assert int(cnt) == want.shape[0] and torch.equal(idx[:want.shape[0]], want)
    mask = v > 0.7
    got, k = rm.compact_mask(mask)
    want = torch.nonzero(mask).squeeze(-1).int()
    assert int(k) == want.shape[0] and torch.equal(got[:want.shape[0]], want)

"""Scripted inputs for the occupancy-grid parity test (TEST INFRASTRUCTURE): a density field made of IEEE-exact torch
operations (bit-identical on CPU and GPU) and hash-based "random" draws that the golden generator feeds to the reference's own
`update_extra_state` (by patching torch.rand_like / torch.randint) and the GPU test feeds to this repo's kernels.
Shared by tests/golden/make_golden_grid_state.py and tests/test_gpu_occupancy.py."""
import numpy as np
import torch

M64 = (1 << 64) - 1


def hash24(key, idx):
    """24-bit hash of (key, idx) — idx a numpy integer array; splitmix64 finaliser"""
    with np.errstate(over="ignore"):
        x = (np.uint64(key) * np.uint64(0x9E3779B97F4A7C15) + idx.astype(np.uint64) * np.uint64(0xBF58476D1CE4E5B9) + np.uint64(0x94D049BB133111EB))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(40)).astype(np.int64)


def uniform(key, idx):
    """float32 in [0,1) with 24 random bits (exactly representable)"""
    return (hash24(key, idx).astype(np.float64) / float(1 << 24)).astype(np.float32)


def scripted_density(x):
    """sigma(x) >= 0 from separate IEEE-exact torch ops (mul, add, sub, clamp): two paraboloid blobs.  x [n,3] fp32."""
    x0, x1, x2 = x[:, 0], x[:, 1], x[:, 2]
    r2 = (x0 * x0) + (x1 * x1) + (x2 * x2)
    a = torch.clamp(18.0 - r2 * 25.0, min=0.0)
    y0, y1, y2 = x0 - 0.9, x1 + 0.35, x2 - 0.2
    q2 = (y0 * y0) + (y1 * y1) + (y2 * y2)
    b = torch.clamp(12.0 - q2 * 60.0, min=0.0)
    return a + b


def full_noise(upd, cas, morton_idx):
    """jitter variates [n,3] of the full refresh number `upd` for the cells `morton_idx` of cascade `cas`"""
    m = np.asarray(morton_idx, np.int64)
    return np.stack([uniform(1000 + upd * 16 + cas, m * 3 + a) for a in range(3)], axis=-1)


def partial_draws(upd, cas, n_pick, H, n_occ):
    """(rand_coords [n_pick,3] int, rand_occ [n_pick] int, noise [2*n_pick,3] fp32) of the partial refresh number `upd`"""
    k = np.arange(n_pick, dtype=np.int64)
    coords = np.stack([(hash24(2000 + upd * 16 + cas, k * 3 + a) * H) >> 24 for a in range(3)], axis=-1)
    occ = (hash24(3000 + upd * 16 + cas, k) * int(n_occ)) >> 24
    k2 = np.arange(2 * n_pick, dtype=np.int64)
    noise = np.stack([uniform(4000 + upd * 16 + cas, k2 * 3 + a) for a in range(3)], axis=-1)
    return coords.astype(np.int64), occ.astype(np.int64), noise


def checksum64(grid):
    """order-sensitive 64-bit checksum of the float bit patterns of a grid (numpy fp32 array)"""
    bits = np.ascontiguousarray(grid, np.float32).reshape(-1).view(np.uint32).astype(np.uint64)
    w = (np.arange(bits.shape[0], dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))
    with np.errstate(over="ignore"):
        return int((bits * w).sum(dtype=np.uint64))


def snapshot(grid, bitfield, stride=127):
    g = np.ascontiguousarray(grid, np.float32).reshape(-1)
    return {"checksum": np.uint64(checksum64(g)), "sum": np.float64(g.astype(np.float64).sum()), "n_neg": np.int64((g < 0).sum()),
            "n_pos": np.int64((g > 0).sum()), "sample": g[::stride].copy(), "bitfield": np.ascontiguousarray(bitfield, np.uint8).copy()}


def camera_ring(n, radius, seed=3):
    """n camera-to-world matrices [n,4,4] on a ring looking at the origin (OpenCV convention: +z forward), slightly tilted"""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4, 4), np.float32)
    for i in range(n):
        ang = 2 * np.pi * i / n + 0.13
        pos = np.array([radius * np.cos(ang), radius * np.sin(ang), 0.35 * radius * np.sin(2.3 * ang)])
        fwd = -pos / np.linalg.norm(pos) + rng.normal(size=3) * 0.05
        fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        out[i, :3, 0], out[i, :3, 1], out[i, :3, 2], out[i, :3, 3] = right, down, fwd, pos
        out[i, 3, 3] = 1.0
    return out

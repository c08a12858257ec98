"""Helpers shared by the `-m gpu` parity tests."""
import numpy as np
import torch

from enerf_b200 import synthetic
from oracle import oracle, ref

DEV = "cuda"


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return x if dtype is None else x.to(dtype)


def n(x):
    return x.detach().cpu().numpy()


def scene(n_rays, bound, seed=0):
    """rays + analytic-ball bitfield + near/far (numpy, via the oracle)."""
    cascade = 1 + int(np.ceil(np.log2(bound)))
    grid = synthetic.ball_density_grid(bound, cascade)
    bits = synthetic.packbits_np(grid)
    o, d = synthetic.random_rays(n_rays, bound, seed=seed)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    return dict(o=o, d=d, aabb=aabb, bits=bits, grid=grid, cascade=cascade, nears=nears, fars=fars, bound=bound)


def per_ray(rays, *arrays):
    """dict ray_id -> tuple of that ray's slices of the [M,...] arrays"""
    out = {}
    for rid, off, cnt in rays:
        out[int(rid)] = tuple(a[off:off + cnt] for a in arrays)
    return out


def ref_mod(name):
    m = ref.load(name)
    return m


def gpu_level_scales(per_level_scale, H, L):
    """exp2f(level*S)*H-1 evaluated on the device, so the oracle can use the GPU's values"""
    S = np.float32(np.log2(per_level_scale))
    lv = torch.arange(L, device=DEV, dtype=torch.float32)
    return n(torch.exp2(lv * float(S)) * float(H) - 1.0).astype(np.float32)

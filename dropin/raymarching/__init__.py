"""Bare-name drop-in for the reference's `raymarching` package: re-exports enerf_b200.raymarching."""
from enerf_b200.raymarching import *  # noqa: F401,F403
from enerf_b200.raymarching import backend as _backend_module  # noqa: F401
from enerf_b200.raymarching.backend import _backend  # noqa: F401
from enerf_b200.raymarching.raymarching import (near_far_from_aabb, polar_from_ray, morton3D, morton3D_invert, packbits,  # noqa: F401
                                                 march_rays_train, composite_rays_train, march_rays, composite_rays, compact_rays,
                                                 composite_uniform)

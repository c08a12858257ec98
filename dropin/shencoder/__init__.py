"""Bare-name drop-in for the reference's `shencoder` package: re-exports enerf_b200.shencoder."""
from enerf_b200.shencoder import *  # noqa: F401,F403
from enerf_b200.shencoder import backend as _backend_module  # noqa: F401
from enerf_b200.shencoder.backend import _backend  # noqa: F401
from enerf_b200.shencoder.sphere_harmonics import SHEncoder, sh_encode, _sh_encoder  # noqa: F401

"""Bare-name drop-in for the reference's `gridencoder` package: re-exports enerf_b200.gridencoder."""
from enerf_b200.gridencoder import *  # noqa: F401,F403
from enerf_b200.gridencoder import backend as _backend_module  # noqa: F401
from enerf_b200.gridencoder.backend import _backend  # noqa: F401
from enerf_b200.gridencoder.grid import GridEncoder, grid_encode, _grid_encode  # noqa: F401

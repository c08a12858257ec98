"""Bare-name drop-in for the reference's `ffmlp` package: re-exports enerf_b200.ffmlp."""
from enerf_b200.ffmlp import *  # noqa: F401,F403
from enerf_b200.ffmlp import backend as _backend_module  # noqa: F401
from enerf_b200.ffmlp.backend import _backend  # noqa: F401
from enerf_b200.ffmlp.ffmlp import FFMLP, ffmlp_forward, _ffmlp_forward, convert_activation  # noqa: F401

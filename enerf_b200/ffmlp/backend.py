"""Reference-compatible name (ffmlp/backend.py:34): the prebuilt C-ABI shim, never a JIT build."""
from ..backends import ffmlp_backend as _backend

__all__ = ['_backend']

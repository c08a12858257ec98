"""Reference-compatible name (gridencoder/backend.py:31): the prebuilt C-ABI shim, never a JIT build."""
from ..backends import gridencoder_backend as _backend

__all__ = ['_backend']

"""Loader for the reference's own CUDA extensions built into oracle/_ref/ (TEST INFRASTRUCTURE).

`load('_raymarching')` returns the pybind11 module compiled by oracle/build_ref.py from the
unmodified reference sources, or None when it has not been built.  Used only by the GPU parity
tests and by tests/golden/make_golden.py; never by the product.
"""
import importlib.machinery
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ("_raymarching", "_gridencoder", "_shencoder", "_ffmlp")
_cache = {}


def path(name):
    return os.path.join(_HERE, "_ref", name, name + ".so")


def available(name):
    return os.path.exists(path(name))


def load(name):
    if name in _cache:
        return _cache[name]
    mod = None
    if available(name):
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        loader = importlib.machinery.ExtensionFileLoader(name, path(name))
        spec = importlib.util.spec_from_file_location(name, path(name), loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    _cache[name] = mod
    return mod

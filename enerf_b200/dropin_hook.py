"""Import hook that makes the UNMODIFIED reference (knelk/enerf) use this repository — no file of the reference is edited.

`python main_nerf.py` puts the reference's root first on `sys.path`, so its own `raymarching/`, `gridencoder/`, `shencoder/`,
`ffmlp/` directories and `nerf/renderer.py`, `nerf/network.py`, `nerf/network_ff.py` would win over anything on PYTHONPATH.
A meta-path finder does not care about the order of `sys.path`: `install()` puts one in front that resolves

    raymarching, gridencoder, shencoder, ffmlp (and their submodules)  ->  enerf_b200.<same name>
    nerf.renderer, nerf.network, nerf.network_ff                        ->  enerf_b200.nerf.<same name>

while everything else of the reference (`nerf.utils`, `nerf.provider`, `nerf.gui`, `main_nerf.py`, `loss.py`, ...) imports as
before.  Ways to switch it on:
    PYTHONPATH=/path/to/enerf-b200/dropin:/path/to/enerf-b200  python main_nerf.py ...      (dropin/sitecustomize.py calls install();
                                                       ENERF_DROPIN=packages keeps the reference's own nerf/renderer.py and networks, =off disables)
    python -m enerf_b200.run_reference main_nerf.py ...                                     (launcher, same effect)
With it `--ff --cuda_ray` (which crashes in the reference at HEAD) and `out_dim_color = 1` with `--cuda_ray` work,
and the shipped configuration (`ff = False`, `cuda_ray = False`) runs its MLPs on the tcgen05 kernels.
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

_PACKAGES = {"raymarching": "enerf_b200.raymarching", "gridencoder": "enerf_b200.gridencoder", "shencoder": "enerf_b200.shencoder",
             "ffmlp": "enerf_b200.ffmlp"}
_MODULES = {"nerf.renderer": "enerf_b200.nerf.renderer", "nerf.network": "enerf_b200.nerf.network", "nerf.network_ff": "enerf_b200.nerf.network_ff"}


def _target(name):
    if name in _MODULES:
        return _MODULES[name]
    head, _, rest = name.partition(".")
    if head in _PACKAGES:
        return _PACKAGES[head] + ("." + rest if rest else "")
    return None


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)          # the alias IS the enerf_b200 module (same object, same classes)

    def exec_module(self, module):
        pass


class DropInFinder(importlib.abc.MetaPathFinder):
    def __init__(self, mirrors=True):
        self.mirrors = mirrors          # False: only the four extension packages are redirected; nerf/*.py stay the reference's files

    def find_spec(self, name, path=None, target=None):
        t = _target(name)
        if t is None or (name in _MODULES and not self.mirrors):
            return None
        try:
            real = importlib.util.find_spec(t)
        except (ImportError, ValueError):
            real = None
        if real is None:
            return None
        spec = importlib.machinery.ModuleSpec(name, _AliasLoader(t), is_package=real.submodule_search_locations is not None)
        return spec


def install(mirrors=True):
    """idempotent; returns the finder.  mirrors=False (or ENERF_DROPIN=packages in the environment of dropin/sitecustomize.py) redirects
    only raymarching / gridencoder / shencoder / ffmlp and leaves nerf/renderer.py, nerf/network*.py to the reference."""
    import importlib.util  # noqa: F401
    for f in sys.meta_path:
        if isinstance(f, DropInFinder):
            f.mirrors = mirrors
            return f
    finder = DropInFinder(mirrors)
    sys.meta_path.insert(0, finder)
    return finder


def uninstall():
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, DropInFinder)]

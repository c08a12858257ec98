"""The reference's OWN Python for this path, importable on the CPU (TEST INFRASTRUCTURE, build container only).

`load()` imports `nerf/renderer.py`, `nerf/network.py`, `encoding.py` and `activation.py` from /root/reference unchanged.
Two things stand between those files and a GPU-less container:
  * third-party imports the container lacks (trimesh, mcubes, ...) — replaced by empty stub modules;
  * the three ops that exist only as CUDA extensions in the reference (`raymarching`, `gridencoder`, `shencoder`) — replaced
    by shim modules with the same names whose functions run the C oracle (oracle/enerf_oracle.c) on CPU tensors.
Everything else (NeRFRenderer.run / update_extra_state / mark_untrained_grid, NeRFNetwork, sample_pdf, trunc_exp) is the
reference's code executing as written.  Used to pin `oracle/cpu_reference.py` and to produce tests/golden/grid_state.npz;
never imported by the product, never needed on the GPU box.
"""
import contextlib
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import cpu_reference, oracle

REF = "/root/reference"
_STUBS = ("trimesh", "mcubes", "tensorboardX", "lpips", "torch_ema", "h5py", "imageio", "configargparse", "dearpygui", "matplotlib",
          "mpl_toolkits", "skimage", "pyvista", "turtle", "tkinter", "hdf5plugin", "kornia", "open3d", "pytorch3d", "ffmlp")
_cache = {}


def available():
    return os.path.isdir(os.path.join(REF, "nerf"))


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """last-resort finder: the listed third-party packages (and their submodules) import as empty stubs"""

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] not in _STUBS:
            return None
        return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


class _GridEncoder(cpu_reference.GridEncoderCPU):
    """constructor signature of gridencoder/grid.py:91-97 on the CPU oracle kernels"""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash'):
        assert gridtype == 'hash' and desired_resolution is not None
        super().__init__(input_dim, num_levels, level_dim, base_resolution, log2_hashmap_size, desired_resolution)


class _SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        assert input_dim == 3 and degree == 4
        self.degree, self.output_dim = degree, degree ** 2

    def forward(self, inputs, size=1):
        return cpu_reference.sh_encode_torch(inputs / size)


def _shims():
    rm = types.ModuleType("raymarching")

    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = oracle.near_far_from_aabb(rays_o.detach().numpy().reshape(-1, 3), rays_d.detach().numpy().reshape(-1, 3), aabb.numpy(), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)

    rm.near_far_from_aabb = near_far_from_aabb
    rm.morton3D = lambda coords: torch.from_numpy(oracle.morton3D(coords.numpy().astype(np.int32)).astype(np.int32))
    rm.morton3D_invert = lambda idx: torch.from_numpy(oracle.morton3D_invert(idx.numpy().astype(np.int32)).astype(np.int32))

    def packbits(grid, thresh, bitfield=None):
        bits = torch.from_numpy(oracle.packbits(grid.detach().numpy().reshape(-1), float(thresh)))
        if bitfield is not None:
            bitfield.copy_(bits)
            return bitfield
        return bits

    rm.packbits = packbits
    ge = types.ModuleType("gridencoder")
    ge.GridEncoder = _GridEncoder
    sh = types.ModuleType("shencoder")
    sh.SHEncoder = _SHEncoder
    return {"raymarching": rm, "gridencoder": ge, "shencoder": sh}


@contextlib.contextmanager
def _patched_imports():
    injected = dict(_shims())
    ours = list(injected) + ["nerf", "nerf.renderer", "nerf.network", "nerf.utils", "encoding", "activation"]
    saved = {k: sys.modules.get(k) for k in ours}
    before = set(sys.modules)
    sys.modules.update(injected)
    finder = _StubFinder()
    sys.meta_path.append(finder)
    sys.path.insert(0, REF)
    try:
        yield
    finally:
        sys.path.remove(REF)
        sys.meta_path.remove(finder)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [m for m in set(sys.modules) - before if isinstance(sys.modules[m], _Stub) or m.split(".")[0] in ("nerf", "utils", "loss")]:
            sys.modules.pop(k, None)


def load():
    """-> namespace with the reference's `NeRFRenderer`, `NeRFNetwork` (nerf/network.py; build it with `make_network`), `sample_pdf`,
    `trunc_exp`"""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("/root/reference is not present (build container only)")
    with _patched_imports():
        for m in ("nerf", "nerf.renderer", "nerf.network", "encoding", "activation"):
            sys.modules.pop(m, None)
        renderer = importlib.import_module("nerf.renderer")
        network = importlib.import_module("nerf.network")
        encoding = importlib.import_module("encoding")
        activation = importlib.import_module("activation")
    def make_network(**kwargs):
        """nerf/network.py's NeRFNetwork(**kwargs); its encoders are imported lazily (encoding.py:57-68), hence the patched context"""
        with _patched_imports():
            return network.NeRFNetwork(**kwargs)

    ns = types.SimpleNamespace(NeRFRenderer=renderer.NeRFRenderer, NeRFNetwork=network.NeRFNetwork, make_network=make_network,
                               sample_pdf=renderer.sample_pdf, trunc_exp=activation.trunc_exp, renderer=renderer, network=network,
                               encoding=encoding)
    _cache["ns"] = ns
    return ns

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/enerf_b200.h declares, and the product path refuses to run without a GPU."""
import pytest
import torch

from enerf_b200 import _lib


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 23
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/enerf_b200.h but not exported"


def test_ctypes_table_matches_header():
    declared = set(_lib.declared_symbols())
    bound = set(_lib._SIGS) | {"enerf_last_error", "enerf_abi_version", "enerf_launch_count"}
    assert declared == bound, declared ^ bound


def test_header_is_plain_c():
    src = open(_lib.HEADER_PATH).read()
    assert 'extern "C"' in src
    includes = [l for l in src.splitlines() if l.strip().startswith("#include")]
    assert includes == ["#include <stdint.h>"], includes   # no torch / ATen / CUDA types at the boundary


def test_abi_version_and_error_string():
    L = _lib.lib()
    assert L.enerf_abi_version() == 8
    assert isinstance(L.enerf_last_error(), bytes)
    assert _lib.launch_count() >= 0


def test_argument_validation_needs_no_gpu():
    # bad arguments are rejected before any CUDA call, with a readable message
    L = _lib.lib()
    rc = L.enerf_sh_encode_forward(None, None, 4, 3, 9, 0, None, 0, None)
    assert rc != 0 and b"degree" in L.enerf_last_error()
    rc = L.enerf_ffmlp_forward(None, None, 128, 32, 16, 48, 2, 0, 6, None, None, None)
    assert rc != 0 and b"hidden_dim" in L.enerf_last_error()
    rc = L.enerf_grid_encode_forward(None, None, None, None, 8, 3, 3, 16, 0.5, 16, 0, None, 0, 0, 0, None)
    assert rc != 0 and b"C must be" in L.enerf_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_path_has_no_cpu_fallback():
    from enerf_b200 import gridencoder, raymarching, shencoder
    enc = gridencoder.GridEncoder(num_levels=2, log2_hashmap_size=10)
    with pytest.raises(RuntimeError):
        enc(torch.rand(8, 3))
    with pytest.raises(RuntimeError):
        shencoder.SHEncoder()(torch.rand(8, 3))
    with pytest.raises((RuntimeError, AssertionError)):
        raymarching.near_far_from_aabb(torch.rand(8, 3), torch.rand(8, 3), torch.tensor([-1., -1, -1, 1, 1, 1]), 0.2)


def test_dropin_packages_expose_reference_names():
    from enerf_b200 import ffmlp, gridencoder, raymarching, shencoder
    for name in ["near_far_from_aabb", "polar_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
                 "composite_rays_train", "march_rays", "composite_rays", "compact_rays"]:
        assert callable(getattr(raymarching, name))
    assert hasattr(gridencoder, "GridEncoder") and hasattr(gridencoder, "grid_encode")
    assert hasattr(shencoder, "SHEncoder") and hasattr(shencoder, "sh_encode")
    assert hasattr(ffmlp, "FFMLP") and hasattr(ffmlp, "ffmlp_forward")
    from enerf_b200 import backends
    for obj, names in [(backends.raymarching_backend, ["packbits", "near_far_from_aabb", "polar_from_ray", "morton3D", "morton3D_invert",
                                                       "march_rays_train", "composite_rays_train_forward", "composite_rays_train_backward",
                                                       "march_rays", "composite_rays", "compact_rays"]),
                       (backends.gridencoder_backend, ["grid_encode_forward", "grid_encode_backward"]),
                       (backends.shencoder_backend, ["sh_encode_forward", "sh_encode_backward"]),
                       (backends.ffmlp_backend, ["ffmlp_forward", "ffmlp_inference", "ffmlp_backward", "allocate_splitk", "free_splitk"])]:
        for n in names:
            assert callable(getattr(obj, n)), n


def test_inference_sample_buffers_clear_only_the_alignment_rows():
    """host logic of raymarching.march_rays: the kernel writes rows [0, n_alive * n_step), the wrapper zeroes the rows the 128-row
    alignment adds and nothing else; the three views share one allocation"""
    from enerf_b200.raymarching.raymarching import _empty_samples, _pad_up
    written = 700 * 26
    M = _pad_up(written, 128)
    assert M % 128 == 0 and 0 < M - written < 128
    xyzs, dirs, deltas = _empty_samples(M, written, torch.device("cpu"))
    assert xyzs.shape == (M, 3) and dirs.shape == (M, 3) and deltas.shape == (M, 2)
    assert xyzs.untyped_storage().data_ptr() == dirs.untyped_storage().data_ptr() == deltas.untyped_storage().data_ptr()
    for buf in (xyzs, dirs, deltas):
        assert buf.is_contiguous() and bool((buf[written:] == 0).all())
    xyzs2, _, _ = _empty_samples(written, written, torch.device("cpu"))       # no alignment rows: nothing to clear
    assert xyzs2.shape == (written, 3)


def test_occupancy_bounds_buffer_layout_matches_the_header():
    """ENERF_OCC_BOUNDS_WORDS in include/enerf_b200.h and the allocation in backends.occupancy_bounds agree"""
    import re
    src = open(_lib.HEADER_PATH).read()
    m = re.search(r"#define ENERF_OCC_BOUNDS_WORDS\(C\) \(\(\(\(6u \* \(C\)\) \+ 3u\) & ~3u\) \+ 8u \* \(C\)\)", src)
    assert m, "the macro changed: update backends.occupancy_bounds and this test"
    import inspect
    from enerf_b200 import backends
    body = inspect.getsource(backends.raymarching_backend.occupancy_bounds)
    assert "((6 * C + 3) & ~3) + 8 * C" in body
    for C in range(1, 17):
        off = (6 * C + 3) & ~3
        assert off % 4 == 0 and off >= 6 * C and off - 6 * C < 4          # float rows 16-byte aligned, right behind the integer rows

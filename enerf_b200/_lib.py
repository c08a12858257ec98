"""ctypes binding of the C-ABI library (include/enerf_b200.h).

The library is the product: if it is missing or cannot be loaded every op raises — there is
no CPU or PyTorch fallback anywhere in this package.
"""
import ctypes as C
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ENERF_B200_LIB") or os.path.join(_HERE, "libenerf_b200.so")     # override: instrumented builds (tools/)
HEADER_PATH = os.path.join(_HERE, "..", "include", "enerf_b200.h")

F32, F16 = 0, 1

_lib = None

_p, _u32, _f32, _int, _u64 = C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.c_uint64

# name -> argtypes (restype is int for all compute entry points)
_SIGS = {
    "enerf_near_far_from_aabb": [_p, _p, _p, _u32, _f32, _p, _p, _p],
    "enerf_polar_from_ray": [_p, _p, _f32, _u32, _p, _p],
    "enerf_morton3D": [_p, _u32, _p, _p],
    "enerf_morton3D_invert": [_p, _u32, _p, _p],
    "enerf_packbits": [_p, _u32, _f32, _p, _p],
    "enerf_march_rays_train": [_p, _p, _p, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _p, _u32, _p],
    "enerf_composite_rays_train_forward": [_p, _p, _p, _p, _u32, _u32, _u32, _p, _p, _p, _p],
    "enerf_composite_rays_train_backward": [_p, _p, _p, _p, _p, _p, _p, _p, _u32, _u32, _u32, _p, _p, _p],
    "enerf_march_rays": [_u32, _u32, _p, _p, _p, _p, _f32, _f32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _u32, _p],
    "enerf_composite_rays": [_u32, _u32, _p, _p, _p, _p, _p, _u32, _p, _p, _p, _p],
    "enerf_compact_rays": [_u32, _p, _p, _p, _p, _p, _p],
    "enerf_march_rays_dev": [_u32, _u32, _p, _p, _p, _p, _f32, _f32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _u32, _p, _p],
    "enerf_composite_rays_dev": [_u32, _u32, _p, _p, _p, _p, _p, _u32, _p, _p, _p, _p, _p],
    "enerf_compact_rays_dev": [_u32, _p, _p, _p, _p, _p, _p, _p],
    "enerf_occupancy_bounds": [_p, _u32, _u32, _p, _p],
    "enerf_march_rays_train_bounded": [_p, _p, _p, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _p, _u32, _p, _p],
    "enerf_march_rays_bounded": [_u32, _u32, _p, _p, _p, _p, _f32, _f32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _u32, _p, _p, _p],
    "enerf_grid_encode_forward": [_p, _p, _p, _p, _u32, _u32, _u32, _u32, _f32, _u32, _int, _p, _u32, _int, _int, _p],
    "enerf_grid_encode_backward": [_p, _p, _p, _p, _p, _u32, _u32, _u32, _u32, _f32, _u32, _int, _p, _p, _u32, _int, _int, _int, _p],
    "enerf_grid_encode_forward_xf": [_p, _f32, _f32, _p, _p, _p, _u32, _u32, _u32, _u32, _f32, _u32, _int, _p, _u32, _int, _int, _p],
    "enerf_grid_encode_backward_xf": [_p, _p, _f32, _f32, _p, _p, _p, _u32, _u32, _u32, _u32, _f32, _u32, _int, _p, _p, _u32, _int, _int, _int, _p],
    "enerf_grid_set_backward_mode": [_int],
    "enerf_grid_set_backward_block": [_int],
    "enerf_grid_set_forward_mode": [_int],
    "enerf_sh_encode_forward": [_p, _p, _u32, _u32, _u32, _int, _p, _int, _p],
    "enerf_sh_encode_backward": [_p, _p, _u32, _u32, _u32, _p, _p, _int, _p],
    "enerf_ffmlp_forward": [_p, _p, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _p, _p, _p],
    "enerf_ffmlp_inference": [_p, _p, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _p, _p, _p],
    "enerf_ffmlp_backward": [_p, _p, _p, _p, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _int, _p, _p, _p, _int, _p, _p],
    "enerf_field_sigma_forward": [_p, _p, _p, _u32, _u32, _p, _p, _p, _p],
    "enerf_field_color_forward": [_p, _p, _u32, _u32, _u32, _p, _p, _p, _p],
    "enerf_field_color_backward": [_p, _p, _u32, _p, _p, _p, _u32, _u32, _p, _p, _p, _p],
    "enerf_field_sigma_backward": [_p, _p, _p, _p, _p, _p, _u32, _u32, _p, _p, _p],
    "enerf_field_infer": [_p, _f32, _f32, _p, _p, _p, _u32, _u32, _f32, _u32, _u32, _p, _u32, _p, _u32, _u32, _u32, _p, _p, _p],
    "enerf_field_infer_alive": [_p, _f32, _f32, _p, _p, _p, _u32, _u32, _f32, _u32, _u32, _p, _u32, _p, _u32, _u32, _u32, _p, _p, _p, _u32, _p],
    "enerf_field_density_forward": [_p, _p, _u32, _u32, _p, _p, _p],
    "enerf_field_density_backward": [_p, _p, _p, _p, _p, _u32, _u32, _p, _p, _p],
    "enerf_field_color_inputs": [_p, _u32, _p, _p, _u32, _u32, _f32, _p, _p, _p],
    "enerf_field_color_inputs_backward": [_p, _p, _u32, _p, _p, _p],
    "enerf_compact_greater": [_p, _f32, _u32, _p, _p, _p, _p],
    "enerf_compact_mask": [_p, _u32, _p, _p, _p, _p],
    "enerf_gather_rows": [_p, _p, _u32, _u32, _u32, _p, _p, _p],
    "enerf_scatter_rows": [_p, _p, _u32, _u32, _p, _p, _p],
    "enerf_weighted_sum_forward": [_p, _p, _u32, _u32, _u32, _p, _p],
    "enerf_weighted_sum_backward": [_p, _p, _p, _u32, _u32, _u32, _p, _p, _p],
    "enerf_occ_points_full": [_p, _u32, _u32, _f32, _p, _u64, _p],
    "enerf_occ_points_partial": [_p, _p, _u32, _u32, _u32, _f32, _p, _p, _p, _p, _p, _u64, _p],
    "enerf_occ_update": [_p, _p, _p, _u32, _u32, _u32, _f32, _f32, _f32, _p, _p, _p, _p, _p],
    "enerf_mark_untrained_grid": [_p, _p, _u32, _f32, _f32, _f32, _f32, _u32, _u32, _f32, _p],
    "enerf_finish_rays_forward": [_p, _p, _p, _p, _p, _p, _int, _f32, _u32, _u32, _p, _p, _p],
    "enerf_finish_rays_backward": [_p, _p, _p, _p, _p, _p, _int, _f32, _u32, _u32, _p, _p, _p],
    "enerf_composite_uniform_forward": [_p, _p, _p, _p, _u32, _u32, _u32, _f32, _p, _p, _p, _p],
    "enerf_composite_uniform_backward": [_p, _p, _p, _p, _p, _p, _p, _u32, _u32, _u32, _f32, _p, _p],
    "enerf_get_rays": [_p, _f32, _f32, _f32, _f32, _u32, _u32, _p, _u32, _u32, _u32, _p, _f32, _p, _p, _p, _p, _p],
    "enerf_event_rays": [_p, _p, _p, _p, _f32, _f32, _f32, _f32, _u32, _p, _f32, _p, _p, _p, _p, _p, _p, _p],
    "enerf_event_loss_forward": [_p, _p, _p, _u32, _u32, _int, _int, _f32, _f32, _f32, _p, _p, _p, _p],
    "enerf_event_loss_backward": [_p, _p, _p, _p, _p, _p, _u32, _u32, _int, _int, _f32, _f32, _f32, _p, _p, _p],
    "enerf_sample_event_pairs": [_p, _p, _p, _p, _u32, _u32, _int, _p, _p, _p, _p, _p, _p, _p, _p],
    "enerf_adam_step": [_p, _p, _int, _p, _p, _u64, _p, _f32, _f32, _f32, _f32, _f32, _p, _p, _f32, _p, _p],
    "enerf_grad_to_half": [_p, _p, _u64, _f32, _p, _p],
    "enerf_ffmlp_set_path": [_int],
    "enerf_ffmlp_set_max_ctas": [_int],
    "enerf_ffmlp_uses_tcgen05": [_u32, _u32, _u32, _u32, _u32],
    "enerf_allocate_splitk": [_u64],
    "enerf_free_splitk": [],
}


def declared_symbols():
    """Every function name include/enerf_b200.h declares."""
    with open(HEADER_PATH) as f:
        src = f.read()
    return sorted(set(re.findall(r"\b(enerf_[A-Za-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m enerf_b200.build` "
                "(enerf_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.enerf_last_error.restype = C.c_char_p
        L.enerf_abi_version.restype = C.c_int
        L.enerf_launch_count.restype = C.c_uint64
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("enerf_b200: " + lib().enerf_last_error().decode())


# ---- optional per-call device timing (bench.py's roofline leg) --------------------------------
_prof = None   # None, or dict name -> list of (start_event, end_event)


def profile_start():
    """Record a CUDA event pair around every C-ABI call from now on (current stream)."""
    global _prof
    _prof = {}


def profile_stop():
    """Stop recording; returns {entry point: (n_calls, total_ms)} after synchronising."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (rec or {}).items()}


_nvtx = False


def enable_nvtx(on=True):
    """Wrap every C-ABI call in an NVTX range named after the entry point (`enerf_grid_encode_forward`, ...), so that profiler
    timelines and `ncu --nvtx --nvtx-include` filters can address the stages of the path by name.  Off by default (a push/pop pair
    costs ~1 us of host time per call)."""
    global _nvtx
    _nvtx = bool(on)


def call(name, *args):
    """Invoke a C-ABI entry point and raise on failure."""
    fn = getattr(lib(), name)
    if _nvtx:
        torch.cuda.nvtx.range_push(name)
        try:
            return _call(fn, name, args)
        finally:
            torch.cuda.nvtx.range_pop()
    return _call(fn, name, args)


def _call(fn, name, args):
    if _prof is None:
        check(fn(*args))
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rc = fn(*args)
    b.record()
    _prof.setdefault(name, []).append((a, b))
    check(rc)


def launch_count():
    return int(lib().enerf_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("enerf_b200: expected a CUDA tensor (there is no CPU path)")


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise RuntimeError(f"enerf_b200: unsupported dtype {t.dtype} (float32 / float16 only)")

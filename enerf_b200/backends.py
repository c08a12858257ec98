"""`_backend` objects with the positional signatures of the reference's four pybind11 modules.

The reference wrappers do `import _raymarching as _backend` (raymarching/raymarching.py:9-12,
gridencoder/grid.py:9-12, shencoder/sphere_harmonics.py:9-12, ffmlp/ffmlp.py:10-13) and call the
20 functions declared in */src/*.h with `at::Tensor`s.  These classes expose the same names and
argument order over torch tensors and forward to the C-ABI (include/enerf_b200.h) on the current
CUDA stream, so code written against the reference extensions runs unchanged.
"""
import torch

from . import _lib
from ._lib import dtype_code, need_cuda, ptr, stream


def _contig(name, *ts):
    for t in ts:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")


class _Raymarching:
    """raymarching/src/raymarching.h:7-18 (bindings: raymarching/src/bindings.cpp:5-19)."""

    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
        need_cuda(rays_o, rays_d, aabb, nears, fars)
        _lib.call("enerf_near_far_from_aabb", ptr(rays_o), ptr(rays_d), ptr(aabb), N, min_near, ptr(nears), ptr(fars), stream())

    @staticmethod
    def polar_from_ray(rays_o, rays_d, radius, N, coords):
        need_cuda(rays_o, rays_d, coords)
        _lib.call("enerf_polar_from_ray", ptr(rays_o), ptr(rays_d), radius, N, ptr(coords), stream())

    @staticmethod
    def morton3D(coords, N, indices):
        need_cuda(coords, indices)
        _lib.call("enerf_morton3D", ptr(coords), N, ptr(indices), stream())

    @staticmethod
    def morton3D_invert(indices, N, coords):
        need_cuda(indices, coords)
        _lib.call("enerf_morton3D_invert", ptr(indices), N, ptr(coords), stream())

    @staticmethod
    def packbits(grid, N, density_thresh, bitfield):
        need_cuda(grid, bitfield)
        _lib.call("enerf_packbits", ptr(grid), N, density_thresh, ptr(bitfield), stream())

    @staticmethod
    def occupancy_bounds(grid, C, H):
        """extension: int32 [C, 6] box around the occupied cells of each cascade level (a view of the buffer enerf_occupancy_bounds fills; pass it
        on as it is), or None when H is not a power of two >= 4"""
        if H < 4 or H & (H - 1) or grid.data_ptr() % 16 or grid.numel() * 8 < C * H ** 3:
            return None                                   # the marcher then probes every candidate, as the reference does
        need_cuda(grid)
        words = ((6 * C + 3) & ~3) + 8 * C                # ENERF_OCC_BOUNDS_WORDS: the integer rows, then the float rows the marchers read
        buf = torch.empty(words, dtype=torch.int32, device=grid.device)
        _lib.call("enerf_occupancy_bounds", ptr(grid), C, H, ptr(buf), stream())
        return buf[:6 * C].view(C, 6)                     # same storage, same address: the float rows stay behind it

    @staticmethod
    def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, perturb,
                         occ_bounds=None):
        """`occ_bounds` (extension): the result of occupancy_bounds() for this bitfield; the march then steps over the candidates outside
        the occupied box without probing the grid — same samples, bit for bit"""
        need_cuda(rays_o, rays_d, grid, nears, fars, xyzs, dirs, deltas, rays, counter, occ_bounds)
        _lib.call("enerf_march_rays_train_bounded", ptr(rays_o), ptr(rays_d), ptr(grid), bound, dt_gamma, max_steps, N, C, H, M,
                                                        ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter),
                                                        int(perturb), ptr(occ_bounds), stream())

    @staticmethod
    def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, weights_sum, depth, image):
        need_cuda(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        n_ch = rgbs.shape[-1] if rgbs.dim() > 1 else 1
        _lib.call("enerf_composite_rays_train_forward", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), M, N, n_ch,
                                                            ptr(weights_sum), ptr(depth), ptr(image), stream())

    @staticmethod
    def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, grad_sigmas, grad_rgbs):
        need_cuda(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, grad_sigmas, grad_rgbs)
        n_ch = rgbs.shape[-1] if rgbs.dim() > 1 else 1
        _lib.call("enerf_composite_rays_train_backward", ptr(grad_weights_sum), ptr(grad_image), ptr(sigmas), ptr(rgbs), ptr(deltas),
                                                             ptr(rays), ptr(weights_sum), ptr(image), M, N, n_ch, ptr(grad_sigmas),
                                                             ptr(grad_rgbs), stream())

    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas, perturb,
                   n_alive_dev=None, occ_bounds=None):
        """`n_alive_dev` (extension, also on the two functions below): int32 device scalar with the true number of alive rays; `n_alive`
        is then an upper bound and the host does not have to read the count back before launching"""
        need_cuda(rays_alive, rays_t, rays_o, rays_d, grid, nears, fars, xyzs, dirs, deltas, n_alive_dev, occ_bounds)
        _lib.call("enerf_march_rays_bounded", n_alive, n_step, ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), bound, dt_gamma,
                                                  max_steps, C, H, ptr(grid), ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas),
                                                  int(perturb), ptr(n_alive_dev), ptr(occ_bounds), stream())

    @staticmethod
    def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, n_alive_dev=None):
        need_cuda(rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, n_alive_dev)
        n_ch = rgbs.shape[-1] if rgbs.dim() > 1 else 1
        _lib.call("enerf_composite_rays_dev", n_alive, n_step, ptr(rays_alive), ptr(rays_t), ptr(sigmas), ptr(rgbs), ptr(deltas), n_ch,
                                                  ptr(weights_sum), ptr(depth), ptr(image), ptr(n_alive_dev), stream())

    @staticmethod
    def compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter, n_alive_dev=None):
        need_cuda(rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter, n_alive_dev)
        _lib.call("enerf_compact_rays_dev", n_alive, ptr(rays_alive), ptr(rays_alive_old), ptr(rays_t), ptr(rays_t_old),
                                                ptr(alive_counter), ptr(n_alive_dev), stream())


class _GridEncoder:
    """gridencoder/src/gridencoder.h:12-13.  Layout [L,B,C] like the reference kernels; the
    *_blc variants use the [B,L*C] layout the Python wrapper hands to the MLP."""

    @staticmethod
    def _check(inputs, embeddings, offsets, *others):
        need_cuda(inputs, embeddings, offsets, *others)
        _contig("inputs", inputs)
        _contig("embeddings", embeddings)
        _contig("offsets", offsets)
        if inputs.dtype != torch.float32:
            raise RuntimeError("inputs must be a float32 tensor")
        if offsets.dtype != torch.int32:
            raise RuntimeError("offsets must be an int tensor")

    @staticmethod
    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx, gridtype, out_layout=0, in_add=0.0,
                            in_mul=0.0):
        """`in_add`, `in_mul` (extension): `inputs` are raw positions and the kernel applies x = (raw + in_add) * in_mul itself"""
        _GridEncoder._check(inputs, embeddings, offsets, outputs, dy_dx)
        _contig("outputs", outputs)
        _lib.call("enerf_grid_encode_forward_xf", ptr(inputs), float(in_add), float(in_mul), ptr(embeddings), ptr(offsets), ptr(outputs), B, D, C, L,
                                                      float(S), H, int(calc_grad_inputs), ptr(dy_dx), gridtype, dtype_code(embeddings), out_layout,
                                                      stream())

    @staticmethod
    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx, grad_inputs, gridtype, out_layout=0,
                             in_add=0.0, in_mul=0.0):
        _GridEncoder._check(inputs, embeddings, offsets, grad, grad_embeddings, dy_dx, grad_inputs)
        _contig("grad", grad)
        _contig("grad_embeddings", grad_embeddings)
        _lib.call("enerf_grid_encode_backward_xf", ptr(grad), ptr(inputs), float(in_add), float(in_mul), ptr(embeddings), ptr(offsets),
                                                       ptr(grad_embeddings), B, D, C, L, float(S), H, int(calc_grad_inputs), ptr(dy_dx),
                                                       ptr(grad_inputs), gridtype, dtype_code(grad), dtype_code(grad_embeddings), out_layout,
                                                       stream())


class _SHEncoder:
    """shencoder/src/shencoder.h:10,13."""

    @staticmethod
    def sh_encode_forward(inputs, outputs, B, D, C, calc_grad_inputs, dy_dx):
        need_cuda(inputs, outputs, dy_dx)
        _contig("inputs", inputs)
        _contig("outputs", outputs)
        _lib.call("enerf_sh_encode_forward", ptr(inputs), ptr(outputs), B, D, C, int(calc_grad_inputs), ptr(dy_dx), dtype_code(inputs), stream())

    @staticmethod
    def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
        need_cuda(grad, inputs, dy_dx, grad_inputs)
        _contig("grad", grad)
        _lib.call("enerf_sh_encode_backward", ptr(grad), ptr(inputs), B, D, C, ptr(dy_dx), ptr(grad_inputs), dtype_code(grad), stream())


class _FFMLP:
    """ffmlp/src/ffmlp.h:8-13."""

    @staticmethod
    def _half(name, *ts):
        for t in ts:
            need_cuda(t)
            if t.dtype != torch.float16:
                raise RuntimeError(f"{name} must be a Half tensor")
            if not t.is_contiguous():
                raise RuntimeError(f"{name} must be a contiguous tensor")

    @staticmethod
    def ffmlp_forward(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, forward_buffer, outputs):
        _FFMLP._half("inputs", inputs)
        _FFMLP._half("weights", weights)
        _lib.call("enerf_ffmlp_forward", ptr(inputs), ptr(weights), B, input_dim, output_dim, hidden_dim, num_layers, activation,
                                             output_activation, ptr(forward_buffer), ptr(outputs), stream())

    @staticmethod
    def ffmlp_inference(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, inference_buffer, outputs):
        _FFMLP._half("inputs", inputs)
        _FFMLP._half("weights", weights)
        _lib.call("enerf_ffmlp_inference", ptr(inputs), ptr(weights), B, input_dim, output_dim, hidden_dim, num_layers, activation,
                                               output_activation, ptr(inference_buffer), ptr(outputs), stream())

    @staticmethod
    def ffmlp_backward(grad, inputs, weights, forward_buffer, B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation,
                       calc_grad_inputs, backward_buffer, grad_inputs, grad_weights):
        _FFMLP._half("grad", grad)
        _FFMLP._half("inputs", inputs)
        _FFMLP._half("weights", weights)
        if forward_buffer is not None:          # None: the tcgen05 kernel recomputes the hidden activations from `inputs`
            _FFMLP._half("forward_buffer", forward_buffer)
        need_cuda(backward_buffer, grad_inputs, grad_weights)
        if grad_weights.dtype == torch.float32:
            scratch = grad_weights
        else:
            scratch = torch.empty(weights.numel(), dtype=torch.float32, device=weights.device)
        _lib.call("enerf_ffmlp_backward", ptr(grad), ptr(inputs), ptr(weights), ptr(forward_buffer), B, input_dim, output_dim, hidden_dim,
                                              num_layers, activation, output_activation, int(calc_grad_inputs), ptr(backward_buffer),
                                              ptr(grad_inputs), ptr(grad_weights), dtype_code(grad_weights), ptr(scratch), stream())

    @staticmethod
    def allocate_splitk(size):
        _lib.call("enerf_allocate_splitk", int(size))

    @staticmethod
    def free_splitk():
        _lib.call("enerf_free_splitk", )


raymarching_backend = _Raymarching()
gridencoder_backend = _GridEncoder()
shencoder_backend = _SHEncoder()
ffmlp_backend = _FFMLP()

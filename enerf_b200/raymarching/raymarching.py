"""Drop-in for the reference's `raymarching` package (raymarching/raymarching.py).

Same public callables, argument order, defaults, dtypes and return shapes:
near_far_from_aabb (:49), polar_from_ray (:80), morton3D (:104), morton3D_invert (:126),
packbits (:155), march_rays_train (:230), composite_rays_train (:286), march_rays (:337),
composite_rays (:362), compact_rays (:382).  Every op runs a sm_100a kernel through the C-ABI
(include/enerf_b200.h) on the current stream; CPU inputs are moved to the GPU exactly where the
reference does so.  Extension over the reference: compositing accepts 1..4 colour channels
(the reference kernels are hard-wired to 3, raymarching/src/raymarching.cu:549-551).
"""
import torch
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .. import _lib
from .backend import _backend

_fwd32 = custom_fwd(device_type='cuda', cast_inputs=torch.float32)
_bwd = custom_bwd(device_type='cuda')


def _rays(t):
    if not t.is_cuda:
        t = t.cuda()
    return t.contiguous().view(-1, 3)


# ---------------------------------------------------------------------------- utils
class _near_far_from_aabb(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """rays_o/rays_d [N,3], aabb [6] -> nears [N], fars [N] (FLT_MAX for rays missing the box)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        if not aabb.is_cuda:
            aabb = aabb.to(rays_o.device)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _backend.near_far_from_aabb(rays_o, rays_d, aabb.contiguous(), N, min_near, nears, fars)
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _polar_from_ray(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, radius):
        """Intersection of each ray with the sphere of `radius` as (theta, phi) in [-1,1]^2."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _backend.polar_from_ray(rays_o, rays_d, radius, N, coords)
        return coords


polar_from_ray = _polar_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """coords int [N,3] in [0,128) -> Morton index int32 [N]."""
        if not coords.is_cuda:
            coords = coords.cuda()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        _backend.morton3D(coords.int().contiguous(), N, indices)
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """Morton index [N] -> coords int32 [N,3]."""
        if not indices.is_cuda:
            indices = indices.cuda()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        _backend.morton3D_invert(indices.int().contiguous(), N, coords)
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, grid, thresh, bitfield=None):
        """grid float [C, H^3] -> uint8 [C*H^3/8]; bit i of byte n is grid.flat[8n+i] > thresh."""
        if not grid.is_cuda:
            grid = grid.cuda()
        grid = grid.contiguous()
        N = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _backend.packbits(grid, N, thresh, bitfield)
        return bitfield


packbits = _packbits.apply


def _pad_up(m, align):
    # the reference always adds a full `align` when m is already aligned (raymarching.py:201-203)
    return m + (align - m % align) if align > 0 else m


def _zero_samples(M, dev):
    """zero-initialised xyzs [M,3], dirs [M,3], deltas [M,2] (raymarching.py:205-207) carved out of one allocation: one fill kernel, not three"""
    buf = torch.zeros(M * 8, dtype=torch.float32, device=dev)
    return buf[:3 * M].view(M, 3), buf[3 * M:6 * M].view(M, 3), buf[6 * M:].view(M, 2)


def _empty_samples(M, written, dev):
    """xyzs [M,3], dirs [M,3], deltas [M,2] carved out of one allocation, rows [written, M) zeroed, the others left to the kernel"""
    buf = torch.empty(M * 8, dtype=torch.float32, device=dev)
    xyzs, dirs, deltas = buf[:3 * M].view(M, 3), buf[3 * M:6 * M].view(M, 3), buf[6 * M:].view(M, 2)
    if M > written:
        xyzs[written:].zero_()
        dirs[written:].zero_()
        deltas[written:].zero_()
    return xyzs, dirs, deltas


# ---------------------------------------------------------------------------- train
class _march_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        """Occupancy-grid marching.  Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3] int32
        (ray id, first sample, sample count); step_counter (int32[2]) += (samples, rays)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        if not density_bitfield.is_cuda:
            density_bitfield = density_bitfield.cuda()
        density_bitfield = density_bitfield.contiguous()
        dev = rays_o.device
        N = rays_o.shape[0]
        exact = force_all_rays or mean_count <= 0
        M = N * max_steps if exact else _pad_up(mean_count, align)

        xyzs, dirs, deltas = _zero_samples(M, dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)

        # the box around the occupied cells (one 5 us pass over the bitfield, recomputed on every call so that it can never be stale):
        # the marcher steps over the candidates outside it without probing the grid
        _backend.march_rays_train(rays_o, rays_d, density_bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars,
                                  xyzs, dirs, deltas, rays, step_counter, perturb, _backend.occupancy_bounds(density_bitfield, C, H))

        if exact:
            # first epochs only: size the outputs to the real sample count (one D2H read)
            m = _pad_up(int(step_counter[0].item()), align)
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, deltas, rays):
        """sigmas [M], rgbs [M,c], deltas [M,2], rays [N,3] -> weights_sum [N], depth [N], image [N,c]."""
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        n_ch = rgbs.shape[1]
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, n_ch, dtype=sigmas.dtype, device=sigmas.device)
        _backend.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, weights_sum, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        ctx.dims = (M, N)
        return weights_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is not propagated (raymarching.py:270)
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        M, N = ctx.dims
        if rgbs.dim() == 2 and rgbs.dtype == sigmas.dtype:        # one zero fill for both gradients
            buf = torch.zeros(M * (1 + rgbs.shape[1]), dtype=sigmas.dtype, device=sigmas.device)
            grad_sigmas, grad_rgbs = buf[:M], buf[M:].view(M, rgbs.shape[1])
        else:
            grad_sigmas, grad_rgbs = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        _backend.composite_rays_train_backward(grad_weights_sum.contiguous(), grad_image.contiguous(), sigmas, rgbs, deltas, rays,
                                               weights_sum, image, M, N, grad_sigmas, grad_rgbs)
        return grad_sigmas, grad_rgbs, None, None


composite_rays_train = _composite_rays_train.apply


class _finish_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, weights_sum, depth, image, nears, fars, bg, bg_scalar):
        N, n_ch = image.shape
        weights_sum, depth, image = weights_sum.contiguous(), depth.contiguous(), image.contiguous()
        image_out, depth_out = torch.empty_like(image), torch.empty_like(depth)
        per_ray = int(bg is not None and bg.dim() == 2)
        _lib.call("enerf_finish_rays_forward", _lib.ptr(weights_sum), _lib.ptr(depth), _lib.ptr(image), _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(bg),
                  per_ray, float(bg_scalar), N, n_ch, _lib.ptr(image_out), _lib.ptr(depth_out), _lib.stream())
        ctx.save_for_backward(depth, nears, fars, bg if bg is not None else depth.new_empty(0))
        ctx.meta = (N, n_ch, per_ray, float(bg_scalar), bg is not None)
        return image_out, depth_out

    @staticmethod
    @_bwd
    def backward(ctx, g_image, g_depth):
        depth, nears, fars, bg = ctx.saved_tensors
        N, n_ch, per_ray, bg_scalar, has_bg = ctx.meta
        g_image = None if g_image is None else g_image.contiguous()
        g_depth = None if g_depth is None else g_depth.contiguous()
        g_ws, g_d = torch.empty_like(depth), torch.empty_like(depth)
        _lib.call("enerf_finish_rays_backward", _lib.ptr(g_image), _lib.ptr(g_depth), _lib.ptr(depth), _lib.ptr(nears), _lib.ptr(fars),
                  _lib.ptr(bg) if has_bg else None, per_ray, bg_scalar, N, n_ch, _lib.ptr(g_ws), _lib.ptr(g_d), _lib.stream())
        return g_ws, g_d, g_image, None, None, None, None


def finish_rays(weights_sum, depth, image, nears, fars, bg_color):
    """The two lines that end NeRFRenderer.run_cuda (renderer.py:397-398):
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color;  depth = clamp(depth - nears, min=0) / (fars - nears)
    as one kernel each way (same operation order and roundings as the ATen expression).  A background that needs a gradient (the
    `bg_radius > 0` model) or an unusual shape keeps the ATen expression."""
    N, n_ch = image.shape
    bg, scalar = None, 0.0
    fused = image.is_cuda and image.dtype == torch.float32 and depth.dtype == torch.float32 and weights_sum.dtype == torch.float32
    if isinstance(bg_color, (int, float)):
        scalar = float(bg_color)
    elif torch.is_tensor(bg_color) and not bg_color.requires_grad and bg_color.is_cuda and bg_color.numel() > 0:
        b = bg_color.detach().float()
        # shapes that broadcast against [N, n_ch] the way the ATen expression does (the caller views the result as [..., n_ch] anyway)
        if b.numel() == 1:
            bg = b.reshape(1).expand(n_ch).contiguous()
        elif b.numel() == n_ch and b.shape[-1] == n_ch:
            bg = b.reshape(n_ch).contiguous()
        elif b.numel() == N * n_ch and b.shape[-1] == n_ch and n_ch > 1 or b.shape == (N, n_ch):
            bg = b.reshape(N, n_ch).contiguous()
        else:
            fused = False
    else:
        fused = False
    if not fused:
        return image + (1 - weights_sum).unsqueeze(-1) * bg_color, torch.clamp(depth - nears, min=0) / (fars - nears)
    return _finish_rays.apply(weights_sum, depth, image, nears, fars, bg, scalar)


# ---------------------------------------------------------------------------- inference
class _march_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                perturb=False, dt_gamma=0, max_steps=1024, n_alive_dev=None, occ_bounds=None):
        """March each alive ray up to n_step samples from rays_t; slots without a sample stay zero.  `n_alive_dev` (extension): int32
        device scalar holding the true alive count, `n_alive` then being an upper bound.  `occ_bounds` (extension): `occupancy_bounds()` of
        this bitfield — candidates outside the occupied box are stepped over unprobed (same samples)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        M = _pad_up(n_alive * n_step, align)
        dev = rays_o.device
        # the kernel writes every row of the n_alive * n_step it is launched for (samples, then zeros); only the rows the alignment adds
        # are cleared here — not 32 bytes per sample and round
        xyzs, dirs, deltas = _empty_samples(M, n_alive * n_step, dev)
        _backend.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, density_bitfield,
                            near, far, xyzs, dirs, deltas, perturb, n_alive_dev, occ_bounds)
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


def occupancy_bounds(density_bitfield, C, H):
    """int32 [C, 6] (min x, y, z, max x, y, z in cells) around the occupied cells of each cascade level; None if H is not a power of two"""
    return _backend.occupancy_bounds(density_bitfield.contiguous(), C, H)


class _composite_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, n_alive_dev=None):
        """In-place accumulation into weights_sum/depth/image; rays_t <- -1 for terminated rays."""
        _backend.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas.contiguous(), rgbs.contiguous(), deltas, weights_sum, depth, image,
                                n_alive_dev)
        return tuple()


composite_rays = _composite_rays.apply


class _compact_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter, n_alive_dev=None):
        """Keep rays with rays_t_old >= 0; alive_counter[0] += number kept."""
        _backend.compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter, n_alive_dev)
        return tuple()


compact_rays = _compact_rays.apply


# ---------------------------------------------------------------------------- fused run() integrator
class _composite_uniform(Function):
    """Fixed-step integrator of NeRFRenderer.run (nerf/renderer.py:230-255) as one kernel:
    (sigmas [N,T], z_vals [N,T], nears [N], fars [N], density_scale, num_steps) -> weights [N,T],
    weights_sum [N], depth [N].  Differentiable in sigmas.  No reference ABI (extension).
    `num_steps` = the COARSE step count that defines the last sample's delta, (far-near)/num_steps
    (renderer.py:177,231 keep it after PDF upsampling has made the rows longer); 0 = T."""

    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, z_vals, nears, fars, density_scale=1.0, num_steps=0):
        from .. import _lib
        sigmas, z_vals = sigmas.contiguous(), z_vals.contiguous()
        nears, fars = nears.contiguous().view(-1), fars.contiguous().view(-1)
        _lib.need_cuda(sigmas, z_vals, nears, fars)
        N, T = sigmas.shape
        weights = torch.empty_like(sigmas)
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        _lib.call("enerf_composite_uniform_forward", _lib.ptr(sigmas), _lib.ptr(z_vals), _lib.ptr(nears), _lib.ptr(fars), N, T,
                                                              int(num_steps), float(density_scale), _lib.ptr(weights), _lib.ptr(weights_sum),
                                                              _lib.ptr(depth), _lib.stream())
        ctx.save_for_backward(sigmas, z_vals, nears, fars)
        ctx.density_scale, ctx.num_steps = float(density_scale), int(num_steps)
        return weights, weights_sum, depth

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights, grad_weights_sum, grad_depth):
        from .. import _lib
        sigmas, z_vals, nears, fars = ctx.saved_tensors
        N, T = sigmas.shape
        gw = None if grad_weights is None else grad_weights.contiguous().float()
        gs = None if grad_weights_sum is None else grad_weights_sum.contiguous().float()
        gd = None if grad_depth is None else grad_depth.contiguous().float()
        grad_sigmas = torch.empty_like(sigmas)
        _lib.call("enerf_composite_uniform_backward", _lib.ptr(gw), _lib.ptr(gs), _lib.ptr(gd), _lib.ptr(sigmas), _lib.ptr(z_vals),
                                                               _lib.ptr(nears), _lib.ptr(fars), N, T, ctx.num_steps, ctx.density_scale,
                                                               _lib.ptr(grad_sigmas), _lib.stream())
        return grad_sigmas, None, None, None, None, None


composite_uniform = _composite_uniform.apply


# ---------------------------------------------------------------------------- run(): image = sum_t w * rgb
class _weighted_sum(Function):
    """image [N,C] = sum_t weights [N,T] * rgbs [N,T,C] (renderer.py:255), one warp per ray; differentiable in both."""

    @staticmethod
    @_fwd32
    def forward(ctx, weights, rgbs):
        from .. import _lib
        weights, rgbs = weights.contiguous(), rgbs.contiguous()
        _lib.need_cuda(weights, rgbs)
        N, T = weights.shape
        n_ch = rgbs.shape[-1]
        image = torch.empty(N, n_ch, dtype=torch.float32, device=weights.device)
        _lib.call("enerf_weighted_sum_forward", _lib.ptr(weights), _lib.ptr(rgbs), N, T, n_ch, _lib.ptr(image), _lib.stream())
        ctx.save_for_backward(weights, rgbs)
        return image

    @staticmethod
    @_bwd
    def backward(ctx, grad_image):
        from .. import _lib
        weights, rgbs = ctx.saved_tensors
        N, T = weights.shape
        n_ch = rgbs.shape[-1]
        g = grad_image.contiguous().float()
        gw = torch.empty_like(weights) if ctx.needs_input_grad[0] else None
        gr = torch.empty_like(rgbs) if ctx.needs_input_grad[1] else None
        _lib.call("enerf_weighted_sum_backward", _lib.ptr(g), _lib.ptr(weights), _lib.ptr(rgbs), N, T, n_ch, _lib.ptr(gw), _lib.ptr(gr), _lib.stream())
        return gw, gr


weighted_sum = _weighted_sum.apply


def compact_mask(mask):
    """Indices (int32, increasing) of the True entries of a flat bool tensor and their count — `torch.nonzero(mask)` as three small
    kernels WITHOUT the host round trip: returns (idx [n] — the first `count` entries are valid —, count int32 [1] on the device).
    Consumers (field.masked_color) read the count from the device (the reference's `x[mask]` synchronises, network.py:180-183)."""
    from .. import _lib
    mask = mask.contiguous().view(-1)
    _lib.need_cuda(mask)
    n = mask.shape[0]
    dev = mask.device
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    blocks = torch.empty(-(-n // 4096) + 1, dtype=torch.int32, device=dev)
    _lib.call("enerf_compact_mask", _lib.ptr(mask.view(torch.uint8)), n, _lib.ptr(idx), _lib.ptr(count), _lib.ptr(blocks), _lib.stream())
    return idx, count

"""The reference's hot path chained on the reference's OWN CUDA kernels (TEST INFRASTRUCTURE, GPU box).

`RefStack` is nerf/network_ff.py:51-73 (hash grid -> FFMLP sigma-net -> trunc_exp -> SH ⊕ geo_feat ⊕ 0 -> FFMLP colour-net ->
sigmoid) driven by the training branch of nerf/renderer.py:281-342 (near/far -> march_rays_train -> field -> composite_rays_train),
with every kernel taken from oracle/_ref — the four unmodified extensions of /root/reference built for sm_100a by
oracle/build_ref.py.  The reference's `--ff` network cannot be constructed at HEAD and its Python wrappers are
not on the GPU box, so the thin autograd wrappers are restated here, each citing the wrapper it follows; the arithmetic that
matters (every CUDA kernel, incl. fp16 accumulation in the MLP and fp16 atomics in the hash-grid backward) is the reference's.
Host-side bookkeeping (density-grid refresh, buffers) comes from this repo's mirror `NeRFRenderer`, whose refresh is pinned
bit-for-bit to the reference's Python by tests/test_gpu_occupancy.py.

Used by tests/hotpath_parity.py (training-level PSNR parity of the tcgen05 stack) and tests/test_gpu_hotpath_parity.py; never by
the product.
"""
import numpy as np
import torch
from torch.autograd import Function

from enerf_b200.gridencoder import GridEncoder
from enerf_b200.nerf.renderer import NeRFRenderer
from . import ref


def available():
    return all(ref.available(n) for n in ref.NAMES)


# which extension modules the wrappers below call: the reference's own build (default) or this repo's pybind-compatible backends
# (enerf_b200/backends.py, the same 20 positional signatures) — the latter is "drop-in level 2" of INTEGRATION.md: the reference's
# unfused Python wrappers kept, only `_backend` swapped
_provider = "reference"


def use_backends(which):
    global _provider
    if which not in ("reference", "ours"):
        raise ValueError(which)
    _provider = which


def _ext(name):
    if _provider == "reference":
        return ref.load(name)
    from enerf_b200 import backends
    return {"_raymarching": backends.raymarching_backend, "_gridencoder": backends.gridencoder_backend, "_shencoder": backends.shencoder_backend,
            "_ffmlp": backends.ffmlp_backend}[name]


class _RefGrid(Function):
    """gridencoder/grid.py:19-88: fp16 table under autocast, outputs [L,B,C] permuted to [B,L*C], fp16 atomics in the backward"""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution):
        R = _ext("_gridencoder")
        inputs = inputs.contiguous()
        table = embeddings.half().contiguous()                 # grid.py:38-39 (re-cast on every call)
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, table.shape[1]
        S = float(np.log2(per_level_scale))
        out = torch.empty(L, B, C, device=inputs.device, dtype=table.dtype)
        dummy = torch.empty(1, device=inputs.device, dtype=table.dtype)
        R.grid_encode_forward(inputs, table, offsets, out, B, D, C, L, S, base_resolution, False, dummy, 0)
        ctx.save_for_backward(inputs, table, offsets)
        ctx.dims = (B, D, C, L, S, base_resolution)
        return out.permute(1, 0, 2).reshape(B, L * C)          # grid.py:52

    @staticmethod
    def backward(ctx, grad):
        R = _ext("_gridencoder")
        inputs, table, offsets = ctx.saved_tensors
        B, D, C, L, S, H = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous().to(table.dtype)      # grid.py:70
        gt = torch.zeros_like(table)                                                  # grid.py:72
        dummy = torch.empty(1, device=grad.device, dtype=table.dtype)
        R.grid_encode_backward(grad, inputs, table, offsets, gt, B, D, C, L, S, H, False, dummy, dummy, 0)
        return None, gt.float(), None, None, None


class _RefSH(Function):
    """shencoder/sphere_harmonics.py:14-58 (inputs cast to half under autocast, no input gradient on this path)"""

    @staticmethod
    def forward(ctx, dirs, degree):
        R = _ext("_shencoder")
        x = dirs.half().contiguous()
        B = x.shape[0]
        out = torch.empty(B, degree ** 2, dtype=x.dtype, device=x.device)
        dummy = torch.empty(1, dtype=x.dtype, device=x.device)
        R.sh_encode_forward(x, out, B, 3, degree, False, dummy)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, None


class _RefFFMLP(Function):
    """ffmlp/ffmlp.py:15-86: fp16 in/out, forward_buffer stored, backward = fused dgrad kernel + split-K CUTLASS weight gradients"""

    @staticmethod
    def forward(ctx, x, weights, in_dim, num_layers, calc_grad_inputs):
        R = _ext("_ffmlp")
        x = x.half().contiguous()
        w = weights.half().contiguous()
        B = x.shape[0]
        out = torch.empty(B, 16, dtype=torch.half, device=x.device)
        fb = torch.empty(num_layers, B, 64, dtype=torch.half, device=x.device)
        R.ffmlp_forward(x, w, B, in_dim, 16, 64, num_layers, 0, 6, fb, out)
        ctx.save_for_backward(x, w, fb)
        ctx.dims = (B, in_dim, num_layers, calc_grad_inputs)
        return out

    @staticmethod
    def backward(ctx, g):
        R = _ext("_ffmlp")
        x, w, fb = ctx.saved_tensors
        B, in_dim, nl, want_dx = ctx.dims
        g = g.half().contiguous()
        bb = torch.zeros(nl, B, 64, dtype=torch.half, device=g.device)               # ffmlp.py:67-73
        gi = torch.zeros(B, in_dim, dtype=torch.half, device=g.device)
        gw = torch.zeros_like(w)
        R.ffmlp_backward(g, x, w, fb, B, in_dim, 16, 64, nl, 0, 6, want_dx, bb, gi, gw)
        return (gi if want_dx else None), gw.float(), None, None, None


class _RefTruncExp(Function):
    """activation.py:5-18"""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


class _RefComposite(Function):
    """raymarching/raymarching.py:233-286 (3 colour channels, hard-wired in the reference kernels)"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays):
        R = _ext("_raymarching")
        sigmas, rgbs = sigmas.float().contiguous(), rgbs.float().contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        ws, dp, im = (torch.empty(N, device=sigmas.device), torch.empty(N, device=sigmas.device), torch.empty(N, 3, device=sigmas.device))
        R.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, ws, dp, im)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, im)
        ctx.dims = (M, N)
        return ws, dp, im

    @staticmethod
    def backward(ctx, g_ws, g_dp, g_im):
        R = _ext("_raymarching")
        sigmas, rgbs, deltas, rays, ws, im = ctx.saved_tensors
        M, N = ctx.dims
        gs, gr = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        R.composite_rays_train_backward(g_ws.contiguous(), g_im.contiguous(), sigmas, rgbs, deltas, rays, ws, im, M, N, gs, gr)
        return gs, gr, None, None


def _pad128(x):
    tail = -x.shape[0] % 128
    return x if tail == 0 else torch.cat([x, x.new_zeros(tail, x.shape[1])])


class RefStack(NeRFRenderer):
    """parameters: `encoder.embeddings`, `w_sigma` (32-64-64-16), `w_color` (32-64-64-64-16): the layouts of GridEncoder / FFMLP"""

    def __init__(self, bound=1, density_scale=1, min_near=0.2, density_thresh=0.01):
        super().__init__(bound, cuda_ray=True, density_scale=density_scale, min_near=min_near, density_thresh=density_thresh, bg_radius=-1)
        self.encoder = GridEncoder(desired_resolution=2048 * bound)
        self.w_sigma = torch.nn.Parameter(torch.zeros(64 * (32 + 64 + 16)))
        self.w_color = torch.nn.Parameter(torch.zeros(64 * (32 + 64 * 2 + 16)))
        _ext("_ffmlp").allocate_splitk(4)

    def _h(self, x):
        e = self.encoder
        feat = _RefGrid.apply((x + self.bound) / (2 * self.bound), e.embeddings, e.offsets, e.per_level_scale, e.base_resolution)
        n = feat.shape[0]
        return _RefFFMLP.apply(_pad128(feat), self.w_sigma, 32, 2, feat.requires_grad)[:n]       # ffmlp.py:161: calc_grad_inputs = inputs.requires_grad

    def density(self, x):
        h = self._h(x)
        return {'sigma': _RefTruncExp.apply(h[:, 0]), 'geo_feat': h[:, 1:]}

    def forward(self, x, d):
        """network_ff.py:51-73"""
        h = self._h(x)
        sigma = _RefTruncExp.apply(h[:, 0])
        sh = _RefSH.apply(d, 4)
        cin = torch.cat([sh, h[:, 1:], torch.zeros_like(h[:, :1])], dim=-1)
        n = cin.shape[0]
        y = _RefFFMLP.apply(_pad128(cin), self.w_color, 32, 3, cin.requires_grad)[:n]
        return sigma, torch.sigmoid(y[:, :3])

    def render_train(self, rays_o, rays_d, bg_color=1, perturb=True, force_all_rays=False, max_steps=1024):
        """training branch of NeRFRenderer.run_cuda (renderer.py:281-342) on the reference kernels"""
        R = _ext("_raymarching")
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        N, dev = rays_o.shape[0], rays_o.device
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        R.near_far_from_aabb(rays_o, rays_d, self.aabb_train, N, self.min_near, nears, fars)
        counter = self.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        exact = force_all_rays or self.mean_count <= 0                       # raymarching.py:195-203
        M = N * max_steps if exact else self.mean_count + (128 - self.mean_count % 128)
        xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        R.march_rays_train(rays_o, rays_d, self.density_bitfield, float(self.bound), 0.0, max_steps, N, self.cascade, self.grid_size, M, nears, fars,
                           xyzs, dirs, deltas, rays, counter, 1 if perturb else 0)
        if exact:                                                            # raymarching.py:218-226
            m = int(counter[0].item())
            m = m + (128 - m % 128)
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        with torch.autocast("cuda", dtype=torch.float16):
            sigmas, rgbs = self(xyzs, dirs)
        sigmas = self.density_scale * sigmas
        ws, depth, image = _RefComposite.apply(sigmas, rgbs, deltas, rays)
        image = image + (1 - ws).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {'image': image, 'depth': depth}

    # ---- inference (renderer.py:344-400) on the reference kernels -------------------------------------------------------------
    def _infer_mlp(self, x, w, num_layers):
        """ffmlp.py:40-42: the inference kernel (no forward_buffer)"""
        R = _ext("_ffmlp")
        x = _pad128(x.half().contiguous())
        B = x.shape[0]
        out = torch.empty(B, 16, dtype=torch.half, device=x.device)
        buf = torch.empty(B, 64, dtype=torch.half, device=x.device)
        R.ffmlp_inference(x, w.detach().half().contiguous(), B, 32, 16, 64, num_layers, 0, 6, buf, out)
        return out

    def _infer_field(self, xyzs, dirs):
        """network_ff.py:51-73 without autograd"""
        R, e = _ext("_gridencoder"), self.encoder
        n = xyzs.shape[0]
        x = ((xyzs + self.bound) / (2 * self.bound)).contiguous()
        table = e.embeddings.detach().half().contiguous()
        L, C = e.offsets.shape[0] - 1, table.shape[1]
        feat = torch.empty(L, n, C, device=x.device, dtype=table.dtype)
        dummy = torch.empty(1, device=x.device, dtype=table.dtype)
        R.grid_encode_forward(x, table, e.offsets, feat, n, 3, C, L, float(np.log2(e.per_level_scale)), e.base_resolution, False, dummy, 0)
        h = self._infer_mlp(feat.permute(1, 0, 2).reshape(n, L * C), self.w_sigma, 2)[:n]
        sigma = torch.exp(h[:, 0].float())
        sh = _RefSH.apply(dirs, 4)
        cin = torch.cat([sh, h[:, 1:], torch.zeros_like(h[:, :1])], dim=-1)
        y = self._infer_mlp(cin, self.w_color, 3)[:n]
        return sigma, torch.sigmoid(y[:, :3])

    @torch.no_grad()
    def render_infer(self, rays_o, rays_d, bg_color=1, perturb=False, dt_gamma=0.0, max_steps=1024):
        R = _ext("_raymarching")
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        N, dev = rays_o.shape[0], rays_o.device
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        R.near_far_from_aabb(rays_o, rays_d, self.aabb_infer, N, self.min_near, nears, fars)
        weights_sum, depth, image = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
        n_alive = N
        alive_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        rays_alive = torch.zeros(2, N, dtype=torch.int32, device=dev)
        rays_t = torch.zeros(2, N, device=dev)
        step = i = 0
        while step < 1024:                                                  # renderer.py:365 (hard-coded)
            if step == 0:
                rays_alive[0] = torch.arange(N, dtype=torch.int32, device=dev)
                rays_t[0] = nears
            else:
                alive_counter.zero_()
                R.compact_rays(n_alive, rays_alive[i % 2], rays_alive[(i + 1) % 2], rays_t[i % 2], rays_t[(i + 1) % 2], alive_counter)
                n_alive = int(alive_counter.item())
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            M = n_alive * n_step
            M += 128 - (M % 128)                                            # raymarching.py:326-327
            xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
            R.march_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], rays_o, rays_d, float(self.bound), dt_gamma, max_steps, self.cascade,
                         self.grid_size, self.density_bitfield, nears, fars, xyzs, dirs, deltas, 1 if perturb else 0)
            sigmas, rgbs = self._infer_field(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            R.composite_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], sigmas.float().contiguous(), rgbs.float().contiguous(), deltas, weights_sum,
                             depth, image)
            step += n_step
            i += 1
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {'image': image, 'depth': depth, 'weights_sum': weights_sum}

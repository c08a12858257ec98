"""CPU port of the reference's pure-PyTorch renderer (TEST / BASELINE INFRASTRUCTURE ONLY).

This is what `bench.py`'s cpu_baseline and `--impl reference` time on the GPU box's host cores:
the path every shipped E-NeRF config runs — `NeRFNetwork` of nerf/network.py (no --ff) driven by
`NeRFRenderer.run` (no --cuda_ray), nerf/renderer.py:150-278 — restated in torch on the CPU.
/root/reference does not exist on the GPU box, so the renderer/network code is re-stated here
(`kind: "port"`); the three ops that exist only as CUDA in the reference (hash grid, SH,
near_far_from_aabb) come from the C oracle (oracle/enerf_oracle.c).  Nothing in enerf_b200/
imports this module.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import oracle


class _GridEncodeCPU(torch.autograd.Function):
    """gridencoder/grid.py:19-88 on the CPU oracle kernels (fp32 table)."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution):
        x = inputs.detach().contiguous().numpy()
        out, _ = oracle.grid_encode_forward(x, embeddings.detach().numpy(), offsets.numpy(), per_level_scale, base_resolution)
        L, B, C = out.shape
        ctx.save_for_backward(inputs, offsets)
        ctx.meta = (per_level_scale, base_resolution, embeddings.shape)
        return torch.from_numpy(np.ascontiguousarray(out.transpose(1, 0, 2)).reshape(B, L * C))

    @staticmethod
    def backward(ctx, grad):
        inputs, offsets = ctx.saved_tensors
        pls, H, shape = ctx.meta
        B = inputs.shape[0]
        L = offsets.shape[0] - 1
        g = np.ascontiguousarray(grad.detach().numpy().reshape(B, L, shape[1]).transpose(1, 0, 2))
        gg = oracle.grid_encode_backward(g, inputs.detach().numpy(), offsets.numpy(), shape[0], shape[1], pls, H)
        return None, torch.from_numpy(gg.astype(np.float32)), None, None, None


class GridEncoderCPU(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048):
        super().__init__()
        self.per_level_scale = oracle.per_level_scale_for(desired_resolution, base_resolution, num_levels)
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        offsets = oracle.grid_offsets(input_dim, num_levels, self.per_level_scale, base_resolution, log2_hashmap_size)
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim).uniform_(-1e-4, 1e-4))

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        return _GridEncodeCPU.apply(inputs.view(-1, 3), self.embeddings, self.offsets, self.per_level_scale, self.base_resolution)


def sh_encode_torch(d):
    """degree-4 real SH, closed forms of shencoder/src/shencoder.cu:51-69"""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return torch.stack([
        torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)], dim=-1)


def sample_pdf_cpu(bins, weights, n_samples, det=False):
    """inverse-CDF sampling of new z values, nerf/renderer.py:12-46.  bins [B,T-1], weights [B,T-2] -> [B,n_samples]"""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if det:
        u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples).expand(list(cdf.shape[:-1]) + [n_samples])
    else:
        u = torch.rand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return b0 + (u - c0) / denom * (b1 - b0)


class NeRFNetworkCPU(nn.Module):
    """nerf/network.py:10-132 (sigma-net 32->64->16, colour-net 31->64->64->C, no bias) +
    NeRFRenderer.run (nerf/renderer.py:150-278), fp32, CPU."""

    def __init__(self, bound=1, out_dim_color=3, num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, min_near=0.2,
                 density_scale=1):
        super().__init__()
        self.bound, self.out_dim_color, self.min_near, self.density_scale = bound, out_dim_color, min_near, density_scale
        self.encoder = GridEncoderCPU(desired_resolution=2048 * bound)
        dims = [32] + [hidden_dim] * (num_layers - 1) + [1 + geo_feat_dim]
        self.sigma_net = nn.ModuleList([nn.Linear(i, o, bias=False) for i, o in zip(dims[:-1], dims[1:])])
        dims = [16 + geo_feat_dim] + [hidden_dim] * (num_layers_color - 1) + [out_dim_color]
        self.color_net = nn.ModuleList([nn.Linear(i, o, bias=False) for i, o in zip(dims[:-1], dims[1:])])
        self.register_buffer('aabb', torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32))

    @staticmethod
    def _mlp(layers, h):
        for l, layer in enumerate(layers):
            h = layer(h)
            if l != len(layers) - 1:
                h = F.relu(h, inplace=True)
        return h

    def density(self, x):
        h = self._mlp(self.sigma_net, self.encoder(x, bound=self.bound))
        return torch.exp(h[..., 0]), h[..., 1:]         # trunc_exp forward (activation.py:10)

    def color(self, d, geo_feat, mask):
        rgbs = torch.zeros(mask.shape[0], self.out_dim_color)
        if not mask.any():
            return rgbs
        h = torch.cat([sh_encode_torch(d[mask]), geo_feat[mask]], dim=-1)
        rgbs[mask] = torch.sigmoid(self._mlp(self.color_net, h))
        return rgbs

    def _weights(self, sigma, z_vals, sample_dist):
        """renderer.py:230-234 (and :203-208): the last delta is (far-near)/num_steps even when the rows have been upsampled"""
        deltas = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], sample_dist * torch.ones_like(z_vals[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * self.density_scale * sigma)
        alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        return alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1], deltas

    def render(self, rays_o, rays_d, num_steps=512, bg_color=1, perturb=False, upsample_steps=0):
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        N = rays_o.shape[0]
        nears, fars = oracle.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), self.aabb.numpy(), self.min_near)
        nears, fars = torch.from_numpy(nears).unsqueeze(-1), torch.from_numpy(fars).unsqueeze(-1)
        z_vals = nears + (fars - nears) * torch.linspace(0.0, 1.0, num_steps).unsqueeze(0)
        sample_dist = (fars - nears) / num_steps
        if perturb:
            z_vals = z_vals + (torch.rand(z_vals.shape) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, self.aabb[:3]), self.aabb[3:])
        sigma, geo_feat = self.density(xyzs.reshape(-1, 3))
        sigma, geo_feat = sigma.view(N, num_steps), geo_feat.view(N, num_steps, -1)
        if upsample_steps > 0:                                   # renderer.py:196-228
            with torch.no_grad():
                w0, deltas = self._weights(sigma, z_vals, sample_dist)
                z_mid = z_vals[..., :-1] + 0.5 * deltas[..., :-1]
                new_z = sample_pdf_cpu(z_mid, w0[:, 1:-1], upsample_steps, det=not self.training).detach()
                new_xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z.unsqueeze(-1)
                new_xyzs = torch.min(torch.max(new_xyzs, self.aabb[:3]), self.aabb[3:])
            new_sigma, new_geo = self.density(new_xyzs.reshape(-1, 3))
            z_vals, z_index = torch.sort(torch.cat([z_vals, new_z], dim=1), dim=1)
            xyzs = torch.cat([xyzs, new_xyzs], dim=1)
            xyzs = torch.gather(xyzs, dim=1, index=z_index.unsqueeze(-1).expand_as(xyzs))
            sigma = torch.gather(torch.cat([sigma, new_sigma.view(N, upsample_steps)], dim=1), dim=1, index=z_index)
            geo_feat = torch.cat([geo_feat, new_geo.view(N, upsample_steps, -1)], dim=1)
            geo_feat = torch.gather(geo_feat, dim=1, index=z_index.unsqueeze(-1).expand_as(geo_feat))
        weights, _ = self._weights(sigma, z_vals, sample_dist)
        mask = weights > 1e-4
        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        geo_feat = geo_feat.reshape(-1, geo_feat.shape[-1])
        rgbs = self.color(dirs.reshape(-1, 3), geo_feat, mask.reshape(-1)).view(N, -1, self.out_dim_color)
        weights_sum = weights.sum(dim=-1)
        depth = torch.sum(weights * ((z_vals - nears) / (fars - nears)).clamp(0, 1), dim=-1)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2) + (1 - weights_sum).unsqueeze(-1) * bg_color
        return {'image': image, 'depth': depth, 'weights_sum': weights_sum}


def time_train_steps(n_rays=256, num_steps=512, bound=3, out_dim_color=1, steps=3, warmup=1, seed=0, threads=None):
    """Times full CPU training steps (render fwd + MSE + backward + Adam) of the port.
    Returns dict(rays_per_s, s_per_step, cores, sample)."""
    import os
    import time
    from enerf_b200 import synthetic
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    torch.manual_seed(seed)
    model = NeRFNetworkCPU(bound=bound, out_dim_color=out_dim_color)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    o, d = synthetic.random_rays(n_rays, bound, seed=seed)
    o, d = torch.from_numpy(o), torch.from_numpy(d)
    target = torch.rand(n_rays, out_dim_color)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = model.render(o, d, num_steps=num_steps, perturb=True)
        loss = F.mse_loss(out['image'], target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    s = float(np.median(times))
    return dict(rays_per_s=n_rays / s, s_per_step=s, cores=threads,
                sample=f"{n_rays} rays x {num_steps} fixed steps/ray (run(), nerf/network.py topology, fp32), fwd+bwd+Adam, median of {steps}")

"""`NeRFRenderer` — host-side mirror of the reference integrator (nerf/renderer.py:86-598).

Same constructor, buffers (`aabb_train`, `aabb_infer`, `density_grid`, `density_bitfield`,
`step_counter`), bookkeeping attributes and methods (`render`, `run`, `run_cuda`,
`update_extra_state`, `mark_untrained_grid`, `reset_extra_state`), so the reference's trainer
(nerf/utils.py) drives it unchanged through `model.render(rays_o, rays_d, staged=..., **vars(opt))`.

Differences (results identical up to fp reassociation):
  * `run()` integrates with ONE fused warp-per-ray kernel (`raymarching.composite_uniform`)
    instead of ~25 ATen kernels over [N,T] temporaries (renderer.py:230-255);
  * any number of colour channels 1..4 works in `run_cuda` (the reference is hard-wired to 3,
    renderer.py:341,354,400, while every E-NeRF config trains 1 channel);
  * the inference loop of `run_cuda` marches more steps per round (`inference_batch_samples`, default 2^23 samples per round;
    0 restores the reference's `n_step <= 8`): same per-ray samples, same image, ~10x fewer host round trips;
  * `render(staged=True)` takes the channel count from `self.out_dim_color` when the subclass
    defines it, else from `kwargs['out_dim_color']`, else 3 (the reference requires the attribute,
    renderer.py:581, and only nerf/network.py sets it).
"""
import math

import torch
import torch.nn as nn

from .. import raymarching


def custom_meshgrid(*args):
    return torch.meshgrid(*args, indexing='ij')


def sample_pdf(bins, weights, n_samples, det=False):
    """Inverse-CDF sampling of new z values (renderer.py:12-46).  bins [B,T], weights [B,T-1]."""
    weights = weights + 1e-5
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    if det:
        u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples, device=weights.device)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    else:
        u = torch.rand(list(cdf.shape[:-1]) + [n_samples], device=weights.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_lo, bin_hi = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)


class NeRFRenderer(nn.Module):
    def __init__(self, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1):
        super().__init__()
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = 128
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        self.bg_radius = bg_radius

        aabb_train = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer('aabb_train', aabb_train)
        self.register_buffer('aabb_infer', aabb_train.clone())

        self.cuda_ray = cuda_ray
        if cuda_ray:
            self.register_buffer('density_grid', torch.zeros([self.cascade, self.grid_size ** 3]))
            self.register_buffer('density_bitfield', torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
            self.mean_density = 0
            self.iter_density = 0
            self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
            self.mean_count = 0
            self.local_step = 0
        # samples per inference round (0 = exactly the reference's n_step policy); see run_cuda
        self.inference_batch_samples = 1 << 23

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def color(self, x, d, mask=None, **kwargs):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    # ------------------------------------------------------------------ fixed-step path
    def run(self, rays_o, rays_d, num_steps=128, upsample_steps=128, bg_color=None, perturb=False, **kwargs):
        """rays [B,N,3] -> {'image' [B,N,C], 'depth' [B,N]} with uniform sampling (renderer.py:150-278)."""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        n_ch = kwargs.get("out_dim_color", getattr(self, "out_dim_color", 3))

        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
        nears, fars = nears.unsqueeze(-1), fars.unsqueeze(-1)

        z_vals = nears + (fars - nears) * torch.linspace(0.0, 1.0, num_steps, device=device).unsqueeze(0)
        sample_dist = (fars - nears) / num_steps
        if perturb:
            z_vals = z_vals + (torch.rand(z_vals.shape, device=device) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])

        density_outputs = self.density(xyzs.reshape(-1, 3))
        for k, v in density_outputs.items():
            density_outputs[k] = v.view(N, num_steps, -1)

        if upsample_steps > 0:
            with torch.no_grad():
                w0, _, _ = raymarching.composite_uniform(density_outputs['sigma'].squeeze(-1), z_vals, nears, fars, self.density_scale, num_steps)
                deltas = z_vals[..., 1:] - z_vals[..., :-1]
                z_mid = z_vals[..., :-1] + 0.5 * deltas
                new_z = sample_pdf(z_mid, w0[:, 1:-1], upsample_steps, det=not self.training).detach()
                new_xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z.unsqueeze(-1)
                new_xyzs = torch.min(torch.max(new_xyzs, aabb[:3]), aabb[3:])
            new_density = self.density(new_xyzs.reshape(-1, 3))
            for k, v in new_density.items():
                new_density[k] = v.view(N, upsample_steps, -1)
            z_vals, z_index = torch.sort(torch.cat([z_vals, new_z], dim=1), dim=1)
            xyzs = torch.cat([xyzs, new_xyzs], dim=1)
            xyzs = torch.gather(xyzs, dim=1, index=z_index.unsqueeze(-1).expand_as(xyzs))
            for k in density_outputs:
                both = torch.cat([density_outputs[k], new_density[k]], dim=1)
                density_outputs[k] = torch.gather(both, dim=1, index=z_index.unsqueeze(-1).expand_as(both))

        # fused: deltas, alphas, transmittance scan, weights, weights_sum, depth
        weights, weights_sum, depth = raymarching.composite_uniform(density_outputs['sigma'].squeeze(-1), z_vals, nears, fars, self.density_scale,
                                                                    num_steps)
        mask = weights > 1e-4

        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        for k, v in density_outputs.items():
            density_outputs[k] = v.view(-1, v.shape[-1])
        rgbs = self.color(xyzs.reshape(-1, 3), dirs.reshape(-1, 3), mask=mask.reshape(-1), **density_outputs)
        rgbs = rgbs.view(N, -1, n_ch)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2)

        if self.bg_radius > 0:
            polar = raymarching.polar_from_ray(rays_o, rays_d, self.bg_radius)
            bg_color = self.background(polar, rays_d.reshape(-1, 3))
        elif bg_color is None:
            bg_color = 1
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        return {'depth': depth.view(*prefix), 'image': image.view(*prefix, n_ch)}

    # ------------------------------------------------------------------ occupancy-grid path
    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024, **kwargs):
        """rays [B,N,3] -> {'image','depth'} with the occupancy-grid marcher (renderer.py:281-406)."""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device

        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer, self.min_near)
        if self.bg_radius > 0:
            polar = raymarching.polar_from_ray(rays_o, rays_d, self.bg_radius)
            bg_color = self.background(polar, rays_d)
        elif bg_color is None:
            bg_color = 1

        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield, self.cascade,
                                                                    self.grid_size, nears, fars, counter, self.mean_count, perturb, 128,
                                                                    force_all_rays, dt_gamma, max_steps)
            sigmas, rgbs = self(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays)
        else:
            n_ch = None
            weights_sum = torch.zeros(N, dtype=torch.float32, device=device)
            depth = torch.zeros(N, dtype=torch.float32, device=device)
            image = None
            n_alive = N
            alive_counter = torch.zeros([1], dtype=torch.int32, device=device)
            rays_alive = torch.zeros(2, n_alive, dtype=torch.int32, device=device)
            rays_t = torch.zeros(2, n_alive, dtype=torch.float32, device=device)
            step, i = 0, 0
            self.last_render_stats = {'samples': 0, 'iterations': 0}    # bookkeeping for bench.py (not in the reference)
            while step < 1024:   # hard-coded in the reference as well (renderer.py:364)
                if step == 0:
                    torch.arange(n_alive, out=rays_alive[0])
                    rays_t[0] = nears
                else:
                    alive_counter.zero_()
                    raymarching.compact_rays(n_alive, rays_alive[i % 2], rays_alive[(i + 1) % 2], rays_t[i % 2], rays_t[(i + 1) % 2], alive_counter)
                    n_alive = alive_counter.item()
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)          # the reference's policy (renderer.py:378)
                if self.inference_batch_samples > 0:
                    # B200: march more steps per round while the batch stays below `inference_batch_samples`; the per-ray sample
                    # sequence and compositing order do not depend on how the steps are grouped into rounds, so the image is the
                    # same — only the number of Python rounds (each with a host sync on n_alive) drops from ~1000 to ~100.
                    n_step = max(n_step, min(self.inference_batch_samples // n_alive, 1024 - step))
                xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], rays_o, rays_d, self.bound,
                                                            self.density_bitfield, self.cascade, self.grid_size, nears, fars, 128, perturb,
                                                            dt_gamma, max_steps)
                sigmas, rgbs = self(xyzs, dirs)
                sigmas = self.density_scale * sigmas
                if image is None:
                    n_ch = rgbs.shape[-1]
                    image = torch.zeros(N, n_ch, dtype=torch.float32, device=device)
                raymarching.composite_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], sigmas, rgbs, deltas, weights_sum, depth, image)
                self.last_render_stats['samples'] += n_alive * n_step
                self.last_render_stats['iterations'] += 1
                step += n_step
                i += 1
            if image is None:
                image = torch.zeros(N, kwargs.get("out_dim_color", getattr(self, "out_dim_color", 3)), dtype=torch.float32, device=device)

        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {'depth': depth.view(*prefix), 'image': image.view(*prefix, image.shape[-1])}

    # ------------------------------------------------------------------ occupancy-grid maintenance
    def _cell_centres(self, coords, cas):
        """world position of density-grid cells `coords` int [n,3] in cascade `cas` (renderer.py:499-506)"""
        bound = min(2 ** cas, self.bound)
        half_grid_size = bound / self.grid_size
        xyzs = 2 * coords.float() / (self.grid_size - 1) - 1
        return xyzs * (bound - half_grid_size), half_grid_size

    def _grid_blocks(self, S):
        """Cells of the 128^3 occupancy grid in S^3 blocks: yields (integer coords [n,3], Morton indices [n])."""
        dev = self.density_grid.device
        axis = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        for xs in axis:
            for ys in axis:
                for zs in axis:
                    coords = torch.stack([g.reshape(-1) for g in custom_meshgrid(xs, ys, zs)], dim=-1)
                    yield coords, raymarching.morton3D(coords).long()

    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """Cells that no training camera sees keep density -1 forever (renderer.py:408-471)."""
        if not self.cuda_ray:
            return
        poses = torch.as_tensor(poses).to(self.density_grid.device)
        fx, fy, cx, cy = intrinsic
        seen_by = torch.zeros_like(self.density_grid)
        for coords, indices in self._grid_blocks(S):
            for cas in range(self.cascade):
                world, half_cell = self._cell_centres(coords, cas)
                for first in range(0, poses.shape[0], S):
                    cams = poses[first:first + S]
                    local = (world.unsqueeze(0) - cams[:, :3, 3].unsqueeze(1)) @ cams[:, :3, :3]       # world -> camera frame
                    depth = local[..., 2]
                    inside = (depth > 0) & (local[..., 0].abs() < cx / fx * depth + half_cell * 2) \
                        & (local[..., 1].abs() < cy / fy * depth + half_cell * 2)
                    seen_by[cas, indices] += inside.sum(0).to(seen_by.dtype)
        self.density_grid[seen_by == 0] = -1

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """EMA-max refresh of density_grid, its bitfield and the mean sample count (renderer.py:474-563)."""
        if not self.cuda_ray:
            return
        dev = self.density_grid.device
        fresh = torch.full_like(self.density_grid, -1.0)

        def sample_density(coords, cas):
            centres, half_cell = self._cell_centres(coords, cas)
            jittered = centres + (torch.rand_like(centres) * 2 - 1) * half_cell
            sigma = self.density(jittered)['sigma'].reshape(-1).detach().float()
            return sigma * (self.density_scale * 0.003383)          # density * nominal step length (renderer.py:512)

        if self.iter_density < 16:                                   # first 16 refreshes: every cell of every cascade
            for coords, indices in self._grid_blocks(S):
                for cas in range(self.cascade):
                    fresh[cas, indices] = sample_density(coords, cas)
        else:                                                        # afterwards: a quarter of the cells at random + as many occupied ones
            n_pick = self.grid_size ** 3 // 4
            for cas in range(self.cascade):
                rand_coords = torch.randint(0, self.grid_size, (n_pick, 3), device=dev)
                rand_idx = raymarching.morton3D(rand_coords).long()
                occupied = torch.nonzero(self.density_grid[cas] > 0).squeeze(-1)
                occ_idx = occupied[torch.randint(0, occupied.shape[0], [n_pick], dtype=torch.long, device=dev)]
                occ_coords = raymarching.morton3D_invert(occ_idx)
                fresh[cas, torch.cat([rand_idx, occ_idx])] = sample_density(torch.cat([rand_coords, occ_coords]), cas)

        both = (self.density_grid >= 0) & (fresh >= 0)
        self.density_grid[both] = torch.maximum(self.density_grid[both] * decay, fresh[both])
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        self.iter_density += 1
        self.density_bitfield = raymarching.packbits(self.density_grid, min(self.mean_density, self.density_thresh), self.density_bitfield)

        counted = min(16, self.local_step)
        if counted > 0:
            self.mean_count = int(self.step_counter[:counted, 0].sum().item() / counted)
        self.local_step = 0

    # ------------------------------------------------------------------ dispatcher
    def render(self, rays_o, rays_d, staged=False, max_ray_batch=4096, **kwargs):
        """rays [B,N,3] -> {'image','depth'}; `staged` chunks rays (never with cuda_ray), renderer.py:566-598."""
        _run = self.run_cuda if self.cuda_ray else self.run
        B, N = rays_o.shape[:2]
        device = rays_o.device
        if staged and not self.cuda_ray:
            n_ch = getattr(self, "out_dim_color", kwargs.get("out_dim_color", 3))
            depth = torch.empty((B, N), device=device)
            image = torch.empty((B, N, n_ch), device=device)
            for b in range(B):
                for head in range(0, N, max_ray_batch):
                    tail = min(head + max_ray_batch, N)
                    part = _run(rays_o[b:b + 1, head:tail], rays_d[b:b + 1, head:tail], **kwargs)
                    depth[b:b + 1, head:tail] = part['depth']
                    image[b:b + 1, head:tail] = part['image']
            return {'depth': depth, 'image': image}
        return _run(rays_o, rays_d, **kwargs)

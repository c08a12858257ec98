"""`NeRFNetwork` (torch-MLP variant, what every shipped E-NeRF config runs) — mirror of
nerf/network.py:10-214.  Same constructor arguments and parameter names (`encoder.embeddings`,
`sigma_net.{l}.weight`, `color_net.{l}.weight`, `bg_net.{l}.weight`), so reference checkpoints load.
Under fp16 autocast with the shipped shapes (hidden 64, 2 + 3 layers, 15 geometry features, degree-4 SH) both nets run on the
tcgen05 kernels of csrc/ffmlp_tc.cu (`enerf_b200.field.density_head` / `masked_color`): sigma-net 32-64-16 with the trunc_exp
head, colour-net 31(+1)-64-64-C on the compacted `weights > 1e-4` samples with SH encoding, concatenation, padding, sigmoid and
the scatter back fused around it.  Other shapes, fp32 and CPU tensors take the reference's own formulation (nn.Linear).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import field
from .. import raymarching
from ..activation import trunc_exp
from ..encoding import get_encoder
from .renderer import NeRFRenderer


def _mlp(dims):
    return nn.ModuleList([nn.Linear(i, o, bias=False) for i, o in zip(dims[:-1], dims[1:])])


def _run_mlp(layers, h):
    for l, layer in enumerate(layers):
        h = layer(h)
        if l != len(layers) - 1:
            h = F.relu(h, inplace=True)
    return h


class NeRFNetwork(NeRFRenderer):
    accepts_ray_dirs = True           # color() takes the broadcast [N,T,3] direction view of NeRFRenderer.run

    def __init__(self, encoding="hashgrid", encoding_dir="sphere_harmonics", encoding_bg="hashgrid", num_layers=2, hidden_dim=64,
                 geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64, num_layers_bg=2, hidden_dim_bg=64, bound=1,
                 disable_view_direction=False, out_dim_color=3, **kwargs):
        super().__init__(bound, **kwargs)
        self.disable_view_direction = disable_view_direction
        self.out_dim_color = out_dim_color
        self.use_tensor_cores = True      # False: always the reference's nn.Linear formulation (used by the parity tests)

        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.geo_feat_dim = geo_feat_dim
        self.encoder, self.in_dim = get_encoder(encoding, desired_resolution=2048 * bound)
        self.sigma_net = _mlp([self.in_dim] + [hidden_dim] * (num_layers - 1) + [1 + geo_feat_dim])

        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color
        self.encoder_dir, self.in_dim_dir = get_encoder(encoding_dir)
        # hidden width of the colour net is `hidden_dim` in the reference too (network.py:66-72)
        self.color_net = _mlp([self.in_dim_dir + geo_feat_dim] + [hidden_dim] * (num_layers_color - 1) + [out_dim_color])

        if self.bg_radius > 0:
            self.num_layers_bg = num_layers_bg
            self.hidden_dim_bg = hidden_dim_bg
            self.encoder_bg, self.in_dim_bg = get_encoder(encoding_bg, input_dim=2, num_levels=4, log2_hashmap_size=19, desired_resolution=2048)
            self.bg_net = _mlp([self.in_dim_bg + self.in_dim_dir] + [hidden_dim_bg] * (num_layers_bg - 1) + [out_dim_color])
        else:
            self.bg_net = None

    def _dir_features(self, d):
        e = self.encoder_dir(d)
        return e * 0 if self.disable_view_direction else e * 1

    def _tc_eligible(self, x):
        """the tcgen05 field applies: CUDA, fp16 autocast, the shipped network shapes"""
        return (self.use_tensor_cores and x.is_cuda and torch.is_autocast_enabled('cuda')
                and field.torch_topology_eligible(self.hidden_dim, self.num_layers, self.num_layers_color, self.in_dim, self.in_dim_dir,
                                                  self.geo_feat_dim, getattr(self.encoder_dir, 'degree', -1), self.out_dim_color))

    def density(self, x):
        if self._tc_eligible(x):
            feat = self.encoder(x, bound=self.bound)
            if feat.dtype == torch.float16 and feat.dim() == 2 and feat.shape[0] > 0:
                rows = feat.shape[0]
                tail = -rows % 128
                if tail:
                    feat = torch.cat([feat, feat.new_zeros(tail, feat.shape[1])])
                sigma, h = field.density_head(feat, field.flat_sigma_weights(self.sigma_net), 1)
                if tail:
                    sigma, h = sigma[:rows], h[:rows]
                # `h` = the 16 raw outputs (geo_feat = h[:, 1:]); color() takes it to build its input rows without a copy
                return {'sigma': sigma, 'geo_feat': h[..., 1:], 'h': h}
            h = _run_mlp(self.sigma_net, feat)
        else:
            h = _run_mlp(self.sigma_net, self.encoder(x, bound=self.bound))
        return {'sigma': trunc_exp(h[..., 0]), 'geo_feat': h[..., 1:]}

    def _grid_density(self, xyzs):
        if self._tc_eligible(xyzs) and xyzs.shape[0] % 128 == 0:
            feat = self.encoder(xyzs, bound=self.bound)
            if feat.dtype == torch.float16:
                return field.density_only(feat, field.flat_sigma_weights(self.sigma_net), 1)
        return super()._grid_density(xyzs)

    def forward(self, x, d):
        """sigma [B] fp32, rgb [B, C] for every sample (network.py:104-132) — the field `run_cuda` queries when a config sets
        `cuda_ray` without `ff`"""
        out = self.density(x)
        if 'h' in out:                                   # tensor-core path: the colour-net on all rows (no mask)
            return out['sigma'], self.color(x, d, mask=None, **out)
        h = torch.cat([self._dir_features(d), out['geo_feat']], dim=-1)
        return out['sigma'], torch.sigmoid(_run_mlp(self.color_net, h))

    def background(self, x, d):
        h = torch.cat([self._dir_features(d), self.encoder_bg(x)], dim=-1)
        return torch.sigmoid(_run_mlp(self.bg_net, h))

    def color(self, x, d, mask=None, geo_feat=None, h=None, **kwargs):
        """rgb [B, C] for the rows `mask` selects (zeros elsewhere), network.py:171-199.  `d`: one direction per row, or
        [N, T, 3] with a ray's direction broadcast over its T samples (no copy).  `h`: the sigma-net outputs from density()."""
        B = x.shape[0]
        if h is not None and h.dim() == 2 and h.shape == (B, 16) and h.dtype == torch.float16 and self._tc_eligible(x):
            if d.dim() == 3 and d.stride(1) == 0:                    # expanded per-ray directions
                dirs, dir_div = d[:, 0, :], d.shape[1]
            else:
                dirs, dir_div = d.reshape(-1, 3), 1
            if mask is None:
                idx, count = torch.arange(B, dtype=torch.int32, device=x.device), None
            else:
                idx, count = raymarching.compact_mask(mask)         # the count stays on the device (no sync, graph-capturable)
            rgbs = field.masked_color(h, dirs, dir_div, idx, count, field.flat_color_weights(self.color_net), self.out_dim_color,
                                      0.0 if self.disable_view_direction else 1.0)
            return rgbs.to(x.dtype)
        d = d.reshape(-1, 3)
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], self.out_dim_color, dtype=x.dtype, device=x.device)
            if not mask.any():
                return rgbs
            d, geo_feat = d[mask], geo_feat[mask]
        h = torch.cat([self._dir_features(d), geo_feat], dim=-1)
        h = torch.sigmoid(_run_mlp(self.color_net, h))
        if mask is not None:
            rgbs[mask] = h.to(rgbs.dtype)
            return rgbs
        return h

    def get_params(self, lr):
        params = [
            {'params': self.encoder.parameters(), 'lr': lr},
            {'params': self.sigma_net.parameters(), 'lr': lr},
            {'params': self.encoder_dir.parameters(), 'lr': lr},
            {'params': self.color_net.parameters(), 'lr': lr},
        ]
        if self.bg_radius > 0:
            params.append({'params': self.encoder_bg.parameters(), 'lr': lr})
            params.append({'params': self.bg_net.parameters(), 'lr': lr})
        return params

"""`NeRFNetwork` (fully-fused variant) — mirror of nerf/network_ff.py:11-147.

sigma-net: hashgrid(32) -> FFMLP 32->64->64->16; sigma = trunc_exp(h[0]), geo_feat = h[1:16].
colour-net: [SH(16) | geo_feat(15) | 0] (32) -> FFMLP 32->64->64->64->16 -> sigmoid -> first C.
Unlike the reference at HEAD (whose ff variant crashes on `out_dim_color` /
`disable_view_direction`), those two arguments are accepted and honoured.
"""
import torch

from .. import field
from ..activation import trunc_exp
from ..encoding import get_encoder
from ..ffmlp import FFMLP
from .renderer import NeRFRenderer


class NeRFNetwork(NeRFRenderer):
    def __init__(self, encoding="hashgrid", encoding_dir="sphere_harmonics", num_layers=2, hidden_dim=64, geo_feat_dim=15,
                 num_layers_color=3, hidden_dim_color=64, bound=1, out_dim_color=3, disable_view_direction=False, **kwargs):
        super().__init__(bound, **kwargs)
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.geo_feat_dim = geo_feat_dim
        self.out_dim_color = out_dim_color
        self.disable_view_direction = disable_view_direction
        self.fuse_field = True      # fused sigma/colour heads (enerf_b200.field) when shapes allow; False = module chain
        self.encoder, self.in_dim = get_encoder(encoding, desired_resolution=2048 * bound)
        self.sigma_net = FFMLP(input_dim=self.in_dim, output_dim=1 + self.geo_feat_dim, hidden_dim=self.hidden_dim, num_layers=self.num_layers)

        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color
        self.encoder_dir, self.in_dim_color = get_encoder(encoding_dir)
        self.in_dim_color += self.geo_feat_dim + 1      # pad to 32 (network_ff.py:42)
        self.color_net = FFMLP(input_dim=self.in_dim_color, output_dim=out_dim_color, hidden_dim=self.hidden_dim_color,
                               num_layers=self.num_layers_color)

    def _color_inputs(self, d, geo_feat):
        d = self.encoder_dir(d)
        if self.disable_view_direction:
            d = d * 0
        pad = torch.zeros_like(geo_feat[..., :1])
        return torch.cat([d, geo_feat, pad], dim=-1)

    def forward(self, x, d):
        # x [N,3] in [-bound,bound], d [N,3] unit -> sigma [N] fp32, rgb [N,C]
        if self.fuse_field and torch.is_autocast_enabled('cuda') and not self.disable_view_direction:
            feat = self.encoder(x, bound=self.bound)
            if field.eligible(feat, d, self.hidden_dim, self.hidden_dim_color, self.in_dim, self.in_dim_color, self.geo_feat_dim,
                              getattr(self.encoder_dir, 'degree', -1), self.out_dim_color, self.sigma_net.activation):
                return field.fused_field(feat, d, self.sigma_net.weights, self.color_net.weights, self.num_layers, self.num_layers_color,
                                         self.out_dim_color, self.training and torch.is_grad_enabled())
            h = self.sigma_net(feat)
        else:
            h = self.sigma_net(self.encoder(x, bound=self.bound))
        sigma = trunc_exp(h[..., 0])
        geo_feat = h[..., 1:]
        rgb = torch.sigmoid(self.color_net(self._color_inputs(d, geo_feat)))
        return sigma, rgb

    def density(self, x):
        h = self.sigma_net(self.encoder(x, bound=self.bound))
        return {'sigma': trunc_exp(h[..., 0]), 'geo_feat': h[..., 1:]}

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        # masked colour query (network_ff.py:92-133): rows outside `mask` stay zero
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], self.out_dim_color, dtype=x.dtype, device=x.device)
            if not mask.any():
                return rgbs
            d, geo_feat = d[mask], geo_feat[mask]
        h = torch.sigmoid(self.color_net(self._color_inputs(d, geo_feat)))
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

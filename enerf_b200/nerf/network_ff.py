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
        self.fuse_infer = True      # under torch.no_grad additionally: encoder + both nets as one kernel (field.fused_infer)
        self.encoder, self.in_dim = get_encoder(encoding, desired_resolution=2048 * bound)
        self.sigma_net = FFMLP(input_dim=self.in_dim, output_dim=1 + self.geo_feat_dim, hidden_dim=self.hidden_dim, num_layers=self.num_layers)

        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color
        self.encoder_dir, self.in_dim_color = get_encoder(encoding_dir)
        self.in_dim_color += self.geo_feat_dim + 1      # pad to 32 (network_ff.py:42)
        self.color_net = FFMLP(input_dim=self.in_dim_color, output_dim=out_dim_color, hidden_dim=self.hidden_dim_color,
                               num_layers=self.num_layers_color)

    def _color_inputs(self, d, geo_feat):
        """[SH(d) (16) | geo_feat (15) | 0] — the 32-wide colour-net input of network_ff.py:64-67"""
        sh = self.encoder_dir(d)
        if self.disable_view_direction:
            sh = sh * 0
        return torch.cat([sh, geo_feat, geo_feat.new_zeros(*geo_feat.shape[:-1], 1)], dim=-1)

    def _sigma_head(self, x):
        h = self.sigma_net(self.encoder(x, bound=self.bound))
        return trunc_exp(h[..., 0]), h[..., 1:]

    def forward(self, x, d):
        """x [N,3] in [-bound, bound], d [N,3] unit  ->  sigma [N] fp32, rgb [N, out_dim_color]"""
        fusable = self.fuse_field and not self.disable_view_direction and torch.is_autocast_enabled('cuda')
        if fusable:
            shapes = (self.hidden_dim, self.hidden_dim_color, self.in_dim, self.in_dim_color, self.geo_feat_dim,
                      getattr(self.encoder_dir, 'degree', -1), self.out_dim_color, self.sigma_net.activation)
            training = self.training and torch.is_grad_enabled()
            if (self.fuse_infer and not torch.is_grad_enabled() and self.num_layers == 2 and self.num_layers_color == 3
                    and isinstance(self.bound, (int, float)) and field.infer_eligible(x, d, self.encoder, *shapes)):
                # rendering: gather + sigma-net + colour-net in one kernel, features and colour-net inputs stay on the SM
                # `_alive_rows`: set by the inference loop of NeRFRenderer.run_cuda around this call (device-side alive count x steps)
                return field.fused_infer(x, d, self.encoder, self.bound, self.sigma_net.weights, self.color_net.weights, self.out_dim_color,
                                         alive=getattr(self, '_alive_rows', None))
            feat = self.encoder(x, bound=self.bound)
            if field.eligible(feat, d, *shapes):
                return field.fused_field(feat, d, self.sigma_net.weights, self.color_net.weights, self.num_layers, self.num_layers_color,
                                         self.out_dim_color, training)
            h = self.sigma_net(feat)
            sigma, geo_feat = trunc_exp(h[..., 0]), h[..., 1:]
        else:
            sigma, geo_feat = self._sigma_head(x)
        return sigma, torch.sigmoid(self.color_net(self._color_inputs(d, geo_feat)))

    def density(self, x):
        sigma, geo_feat = self._sigma_head(x)
        return {'sigma': sigma, 'geo_feat': geo_feat}

    def _grid_density(self, xyzs):
        """density alone for the occupancy-grid refresh: gather + sigma-net with the exp head, nothing else written (the generic
        path would store forward_buffer — 1.6 GB for the 6.3 M cells of a bound-3 grid — because the module is in training mode)"""
        if (torch.is_autocast_enabled('cuda') and xyzs.is_cuda and xyzs.shape[0] % 128 == 0 and self.hidden_dim == 64 and self.in_dim == 32
                and self.num_layers == 2 and self.geo_feat_dim == 15 and self.sigma_net.activation == 0):
            feat = self.encoder(xyzs, bound=self.bound)
            if feat.dtype == torch.float16:
                return field.density_only(feat, self.sigma_net.weights, 2)
        return super()._grid_density(xyzs)

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        """Colour query for the rows selected by `mask` (all rows without one); unselected rows stay zero (network_ff.py:92-133)."""
        if mask is None:
            return torch.sigmoid(self.color_net(self._color_inputs(d, geo_feat)))
        rgbs = torch.zeros(mask.shape[0], self.out_dim_color, dtype=x.dtype, device=x.device)
        if mask.any():
            picked = torch.sigmoid(self.color_net(self._color_inputs(d[mask], geo_feat[mask])))
            rgbs[mask] = picked.to(rgbs.dtype)
        return rgbs

    def get_params(self, lr):
        groups = [self.encoder, self.sigma_net, self.encoder_dir, self.color_net]
        if self.bg_radius > 0:
            groups += [self.encoder_bg, self.bg_net]
        return [{'params': m.parameters(), 'lr': lr} for m in groups]

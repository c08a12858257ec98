"""`get_encoder(...)` -> (module, output_dim): the factory the reference's networks call (encoding.py:45-76), same keyword
arguments and encoding names.  'hashgrid' / 'tiledgrid' / 'sphere_harmonics' map to this repository's CUDA encoders,
'frequency' to the classic sinusoidal encoding in plain torch (never on the configured path), 'None' to the identity.
"""
import torch
from torch import nn


class FreqEncoder(nn.Module):
    """gamma(x) = [x, sin(f_1 x), cos(f_1 x), ...] (reference: encoding.py:5-43)."""

    def __init__(self, input_dim, max_freq_log2, N_freqs, log_sampling=True, include_input=True, periodic_fns=(torch.sin, torch.cos)):
        super().__init__()
        self.input_dim, self.include_input, self.periodic_fns = input_dim, include_input, periodic_fns
        if log_sampling:
            bands = torch.linspace(0.0, max_freq_log2, N_freqs).exp2()
        else:
            bands = torch.linspace(1.0, 2.0 ** max_freq_log2, N_freqs)
        self.freq_bands = bands.tolist()
        self.output_dim = input_dim * (int(include_input) + N_freqs * len(periodic_fns))

    def forward(self, input, **kwargs):
        parts = [fn(input * f) for f in self.freq_bands for fn in self.periodic_fns]
        return torch.cat(([input] if self.include_input else []) + parts, dim=-1)


def _identity(x, **kwargs):
    return x


def get_encoder(encoding, input_dim=3, multires=6, degree=4, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=2048, **kwargs):
    if encoding == 'None':
        return _identity, input_dim
    if encoding in ('hashgrid', 'tiledgrid'):
        from .gridencoder import GridEncoder
        enc = GridEncoder(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                          log2_hashmap_size=log2_hashmap_size, desired_resolution=desired_resolution,
                          gridtype={'hashgrid': 'hash', 'tiledgrid': 'tiled'}[encoding])
    elif encoding == 'sphere_harmonics':
        from .shencoder import SHEncoder
        enc = SHEncoder(input_dim=input_dim, degree=degree)
    elif encoding == 'frequency':
        enc = FreqEncoder(input_dim=input_dim, max_freq_log2=multires - 1, N_freqs=multires, log_sampling=True)
    else:
        raise NotImplementedError(f"unknown encoding {encoding!r}")
    return enc, enc.output_dim

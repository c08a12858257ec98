"""Host side of the spherical-harmonics direction encoder — the counterpart of the reference's
shencoder/sphere_harmonics.py, with its public names: `SHEncoder(input_dim=3, degree=4)`, `sh_encode(inputs, degree,
calc_grad_inputs)`.

Semantics kept from the reference: under autocast the directions are converted to half before encoding (:16); the result
has degree^2 channels; a gradient w.r.t. the directions is produced only when the caller's tensor requires one (:46-54),
through the analytic Jacobian the forward kernel writes.
"""
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd.function import Function, once_differentiable

from .backend import _backend


def _scratch(like, *shape):
    return torch.empty(*shape, dtype=like.dtype, device=like.device)


class _sh_encoder(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.half)
    def forward(ctx, dirs, degree, want_jacobian=False):
        dirs = dirs.contiguous()
        n, dim = dirs.shape
        n_out = degree * degree
        encoded = _scratch(dirs, n, n_out)
        jac = _scratch(dirs, n, dim * n_out) if want_jacobian else _scratch(dirs, 1)
        _backend.sh_encode_forward(dirs, encoded, n, dim, degree, want_jacobian, jac)
        ctx.shape3 = (n, dim, degree)
        ctx.want_jacobian = want_jacobian
        ctx.save_for_backward(dirs, jac)
        return encoded

    @staticmethod
    @once_differentiable
    @custom_bwd(device_type='cuda')
    def backward(ctx, upstream):
        if ctx.want_jacobian:
            dirs, jac = ctx.saved_tensors
            n, dim, degree = ctx.shape3
            upstream = upstream.contiguous().to(dirs.dtype)
            d_dirs = torch.zeros_like(dirs)                       # the kernel accumulates (shencoder.cu:378)
            _backend.sh_encode_backward(upstream, dirs, n, dim, degree, jac, d_dirs)
            return d_dirs, None, None
        return None, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        assert input_dim == 3, "SH encoder only support input dim == 3"
        assert 0 < degree <= 8, "SH encoder only supports degree in [1, 8]"
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree * degree

    def extra_repr(self):
        return f"input_dim={self.input_dim}, degree={self.degree}"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        """inputs [..., 3] within [-size, size]  ->  [..., degree^2]"""
        lead = inputs.shape[:-1]
        flat = (inputs / size).reshape(-1, self.input_dim)
        return sh_encode(flat, self.degree, flat.requires_grad).reshape(*lead, self.output_dim)

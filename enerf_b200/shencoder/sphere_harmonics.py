"""Drop-in for the reference's `shencoder` package (shencoder/sphere_harmonics.py).

`SHEncoder(input_dim=3, degree=4)` / `sh_encode(inputs, degree, calc_grad_inputs)` with the
reference's semantics: inputs are cast to half under autocast (sphere_harmonics.py:16), output
[..., degree^2], gradient w.r.t. the directions only when they require grad (:46-54).
"""
import torch
import torch.nn as nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .backend import _backend


class _sh_encoder(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.half)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        inputs = inputs.contiguous()
        B, input_dim = inputs.shape
        output_dim = degree ** 2
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        if calc_grad_inputs:
            dy_dx = torch.empty(B, input_dim * output_dim, dtype=inputs.dtype, device=inputs.device)
        else:
            dy_dx = torch.empty(1, dtype=inputs.dtype, device=inputs.device)
        _backend.sh_encode_forward(inputs, outputs, B, input_dim, degree, calc_grad_inputs, dy_dx)
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims = [B, input_dim, degree]
        ctx.calc_grad_inputs = calc_grad_inputs
        return outputs

    @staticmethod
    @once_differentiable
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        if not ctx.calc_grad_inputs:
            return None, None, None
        inputs, dy_dx = ctx.saved_tensors
        B, input_dim, degree = ctx.dims
        grad = grad.contiguous()
        if grad.dtype != inputs.dtype:
            grad = grad.to(inputs.dtype)
        grad_inputs = torch.zeros_like(inputs)
        _backend.sh_encode_backward(grad, inputs, B, input_dim, degree, dy_dx, grad_inputs)
        return grad_inputs, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        # inputs [..., 3] in [-size, size] -> [..., degree^2]
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, inputs.requires_grad)
        return outputs.reshape(prefix_shape + [self.output_dim])

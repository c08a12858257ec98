"""Drop-in for the reference's `ffmlp` package (ffmlp/ffmlp.py).

`FFMLP(input_dim, output_dim, hidden_dim, num_layers, activation='relu')` keeps the reference's
single flat fp32 parameter `weights` (same segment order and seed-42 initialisation,
ffmlp.py:118-144) so checkpoints are interchangeable, and `ffmlp_forward(...)` keeps the
reference's 10 positional arguments (ffmlp.py:15-86).

Differences underneath: fp32 accumulation (the reference accumulates in fp16); the batch is
padded to a multiple of 128 only when it is not one already (the reference always appends a
block, ffmlp.py:157-159 — results are identical because the padding rows are sliced off);
weight gradients come back in fp32 straight from the fused reduction; no side streams.
"""
import math

import torch
import torch.nn as nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .backend import _backend


class _ffmlp_forward(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.half)
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation,
                inference=False, calc_grad_inputs=False):
        B = inputs.shape[0]
        inputs = inputs.contiguous()
        weights = weights.contiguous()
        if inputs.dtype != torch.half:   # called outside autocast: the kernels are fp16-in
            inputs = inputs.half()
        if weights.dtype != torch.half:
            weights = weights.half()
        outputs = torch.empty(B, output_dim, device=inputs.device, dtype=inputs.dtype)
        if not inference:
            forward_buffer = torch.empty(num_layers, B, hidden_dim, device=inputs.device, dtype=inputs.dtype)
            _backend.ffmlp_forward(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                                   output_activation, forward_buffer, outputs)
            ctx.save_for_backward(inputs, weights, forward_buffer)
            ctx.dims = (input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs)
        else:
            _backend.ffmlp_inference(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                                     output_activation, None, outputs)
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        B = grad.shape[0]
        grad = grad.contiguous()
        if grad.dtype != torch.half:
            grad = grad.half()
        inputs, weights, forward_buffer = ctx.saved_tensors
        input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs = ctx.dims

        if calc_grad_inputs:
            grad_inputs = torch.empty_like(inputs)
        else:
            grad_inputs = torch.empty(1, device=grad.device, dtype=grad.dtype)
        grad_weights = torch.empty(weights.numel(), device=grad.device, dtype=torch.float32)
        # the tcgen05 kernels keep the activation gradients on the SM; only the generic path needs the buffer
        from .. import _lib
        if _lib.lib().enerf_ffmlp_uses_tcgen05(input_dim, hidden_dim, num_layers, activation, output_activation):
            backward_buffer = None
        else:
            backward_buffer = torch.empty(num_layers, B, hidden_dim, device=grad.device, dtype=grad.dtype)

        _backend.ffmlp_backward(grad, inputs, weights, forward_buffer, B, input_dim, output_dim, hidden_dim, num_layers,
                                activation, output_activation, calc_grad_inputs, backward_buffer, grad_inputs, grad_weights)
        gi = grad_inputs if calc_grad_inputs else None
        return gi, grad_weights, None, None, None, None, None, None, None, None


ffmlp_forward = _ffmlp_forward.apply


def convert_activation(act):
    return {'relu': 0, 'exponential': 1, 'sine': 2, 'sigmoid': 3, 'squareplus': 4, 'softplus': 5}.get(act, 6)


class FFMLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation='relu'):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.activation = convert_activation(activation)
        self.output_activation = convert_activation('none')
        self.tensorcore_width = 16

        assert hidden_dim in [16, 32, 64, 128, 256], f"FFMLP only support hidden_dim in [16, 32, 64, 128, 256], but got {hidden_dim}"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP input_dim should be 16 * m (m  > 0), but got {input_dim}"
        assert output_dim <= 16, f"FFMLP current only supports output dim <= 16, but got {output_dim}"
        assert num_layers >= 2, f"FFMLP num_layers should be larger than 2 (3 matmuls), but got {num_layers}"

        self.padded_output_dim = int(math.ceil(output_dim / 16)) * 16
        # one flat parameter: [hidden,input] + (num_layers-1) x [hidden,hidden] + [16,hidden]
        self.num_parameters = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()
        _backend.allocate_splitk(self.num_layers + 1)

    def cleanup(self):
        _backend.free_splitk()

    def __repr__(self):
        return (f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} "
                f"num_layers={self.num_layers} activation={self.activation}")

    def reset_parameters(self):
        torch.manual_seed(42)   # the reference reseeds the global RNG here (ffmlp.py:142)
        std = math.sqrt(3 / self.hidden_dim)
        self.weights.data.uniform_(-std, std)

    def forward(self, inputs):
        # inputs [B, input_dim] -> [B, output_dim]
        B, C = inputs.shape
        pad = (-B) % 128
        if pad > 0:
            inputs = torch.cat([inputs, torch.zeros(pad, C, dtype=inputs.dtype, device=inputs.device)], dim=0)
        outputs = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim, self.num_layers,
                                self.activation, self.output_activation, not self.training, inputs.requires_grad)
        if B != outputs.shape[0] or self.padded_output_dim != self.output_dim:
            outputs = outputs[:B, :self.output_dim]
        return outputs

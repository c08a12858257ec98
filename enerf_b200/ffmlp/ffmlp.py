"""Host side of the fully-fused MLP — the counterpart of the reference's ffmlp/ffmlp.py with its public names:
`FFMLP(input_dim, output_dim, hidden_dim, num_layers, activation='relu')`, `ffmlp_forward(...)` (10 positional arguments,
ffmlp.py:15-86), `convert_activation`.

Kept from the reference so checkpoints and callers are interchangeable: ONE flat fp32 parameter `weights` holding, in this
order, the row-major [hidden, input] first matrix, `num_layers - 1` [hidden, hidden] matrices and the [16, hidden] output
matrix (output width padded to 16); initialisation U(-sqrt(3/hidden), sqrt(3/hidden)) after `torch.manual_seed(42)`
(ffmlp.py:141-144 — yes, it reseeds the global generator); `inference = not self.training`; input gradients only when
the input requires them; everything half under autocast.

Different underneath: fp32 accumulation in the kernels (the reference accumulates in fp16); the batch is padded to a
multiple of 128 only when needed (the reference always appends a block, ffmlp.py:157-159 — the padding rows are sliced
off either way); weight gradients arrive in fp32 from the fused reduction; no side streams, no split-K workspaces.
"""
import math

import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from .backend import _backend

_ACTIVATIONS = ('relu', 'exponential', 'sine', 'sigmoid', 'squareplus', 'softplus')      # codes 0..5; anything else -> 6 ("none")
_TILE = 128                                                                              # rows per kernel tile


def convert_activation(act):
    return _ACTIVATIONS.index(act) if act in _ACTIVATIONS else 6


def _as_half(t):
    t = t.contiguous()
    return t if t.dtype == torch.half else t.half()


class _ffmlp_forward(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.half)
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation,
                inference=False, calc_grad_inputs=False):
        x, w = _as_half(inputs), _as_half(weights)        # also when called outside autocast: the kernels are fp16-in
        rows = x.shape[0]
        y = x.new_empty(rows, output_dim)
        shape = (rows, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation)
        if inference:
            _backend.ffmlp_inference(x, w, *shape, None, y)
            return y
        hidden = x.new_empty(num_layers, rows, hidden_dim)            # post-activation outputs of every hidden matmul
        _backend.ffmlp_forward(x, w, *shape, hidden, y)
        ctx.shape = shape
        ctx.want_dx = calc_grad_inputs
        ctx.save_for_backward(x, w, hidden)
        return y

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, dy):
        from .. import _lib
        x, w, hidden = ctx.saved_tensors
        rows, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation = ctx.shape
        dy = _as_half(dy)
        dx = torch.empty_like(x) if ctx.want_dx else dy.new_empty(1)
        dw = torch.empty(w.numel(), dtype=torch.float32, device=dy.device)
        # the tcgen05 kernels keep the activation gradients on the SM; only the generic mma.sync path needs a buffer for them
        on_sm = _lib.lib().enerf_ffmlp_uses_tcgen05(input_dim, hidden_dim, num_layers, activation, output_activation)
        scratch = None if on_sm else dy.new_empty(num_layers, rows, hidden_dim)
        _backend.ffmlp_backward(dy, x, w, hidden, *ctx.shape, ctx.want_dx, scratch, dx, dw)
        return (dx if ctx.want_dx else None, dw) + (None,) * 8


ffmlp_forward = _ffmlp_forward.apply


class FFMLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation='relu'):
        super().__init__()
        assert hidden_dim in [16, 32, 64, 128, 256], f"FFMLP only support hidden_dim in [16, 32, 64, 128, 256], but got {hidden_dim}"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP input_dim should be 16 * m (m  > 0), but got {input_dim}"
        assert output_dim <= 16, f"FFMLP current only supports output dim <= 16, but got {output_dim}"
        assert num_layers >= 2, f"FFMLP num_layers should be larger than 2 (3 matmuls), but got {num_layers}"
        self.input_dim, self.output_dim, self.hidden_dim, self.num_layers = input_dim, output_dim, hidden_dim, num_layers
        self.activation, self.output_activation = convert_activation(activation), convert_activation('none')
        self.tensorcore_width = 16
        self.padded_output_dim = 16 * math.ceil(output_dim / 16)
        self.num_parameters = hidden_dim * (input_dim + (num_layers - 1) * hidden_dim + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()
        _backend.allocate_splitk(num_layers + 1)          # kept for the reference's call sequence; nothing is allocated

    def reset_parameters(self):
        torch.manual_seed(42)
        bound = math.sqrt(3 / self.hidden_dim)
        with torch.no_grad():
            self.weights.uniform_(-bound, bound)

    def cleanup(self):
        _backend.free_splitk()

    def __repr__(self):
        return (f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} "
                f"num_layers={self.num_layers} activation={self.activation}")

    def forward(self, inputs):
        """[B, input_dim] -> [B, output_dim]"""
        rows = inputs.shape[0]
        tail = -rows % _TILE
        if tail:
            inputs = torch.cat([inputs, inputs.new_zeros(tail, inputs.shape[1])])
        y = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim, self.num_layers,
                          self.activation, self.output_activation, not self.training, inputs.requires_grad)
        if tail or self.padded_output_dim != self.output_dim:
            y = y[:rows, :self.output_dim]
        return y

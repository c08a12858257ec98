"""`trunc_exp` (reference: activation.py:5-18).

Forward is a plain fp32 exponential (autocast inputs are promoted); the derivative is evaluated on the argument clamped
to [-15, 15] so that a huge pre-activation cannot produce an infinite gradient.  In the fused field the same rule lives in
the sigma-net prologue of the backward kernel (csrc/ffmlp_tc.cu, PRO == 2); this stand-alone op serves the module chain.
"""
import torch
from torch.amp import custom_bwd, custom_fwd

_CLAMP = 15.0


class _trunc_exp(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, pre_activation):
        ctx.save_for_backward(pre_activation)
        return pre_activation.exp()

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, upstream):
        pre_activation, = ctx.saved_tensors
        return upstream * pre_activation.clamp(min=-_CLAMP, max=_CLAMP).exp()


trunc_exp = _trunc_exp.apply

"""`FusedAdam` — drop-in for the optimizer E-NeRF builds in main_nerf.py:211-214,
`torch.optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15)`.

Same update rule and state names (`step`, `exp_avg`, `exp_avg_sq`) as torch's Adam, one kernel per parameter tensor
(28 bytes of HBM traffic per parameter), device-side step counters, and the GradScaler protocol of torch's fused optimizers
(`_step_supports_amp_scaling`: `GradScaler.step` hands over `grad_scale` / `found_inf`, the kernel un-scales on the fly and
skips the step on overflow), so a whole training iteration stays capturable in a CUDA graph.  CUDA fp32 parameters only.

The update goes through the parameter's raw device pointer, so its autograd version counter does not move.  Consumers that
cache derived data per version must be told another way: for a hash-grid table with an fp16 shadow
(`enerf_b200.gridencoder.grid.half_shadow`) the kernel writes the shadow in the same pass, so the next forward reads the updated
table without the 52 MB -> 26 MB cast of gridencoder/grid.py:38-39.
"""
import torch

from . import _lib
from ._lib import ptr, stream


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True
    honours_enerf_shard = True        # parallel.ShardedExchange: `param._enerf_shard = (lo, hi, summed gradient slice, 1/world)`

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("FusedAdam does not take a closure")
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        work = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: contiguous CUDA fp32 parameters only (there is no CPU path)")
                # data-parallel sharding (enerf_b200.parallel.ShardedExchange): this rank owns elements [lo, hi) of the flattened
                # parameter, `grad` is the sum over ranks of that slice and `mul` = 1/world undoes the sum
                shard = getattr(p, "_enerf_shard", None)
                if shard is not None:
                    lo, hi, g, mul = shard
                    pv = p.view(-1)[lo:hi]
                else:
                    lo, hi, mul = 0, p.numel(), 1.0
                    pv = p
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    if g.dtype != torch.float32:
                        g = g.float()
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(pv, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(pv, memory_format=torch.preserve_format)
                if st["exp_avg"].numel() != hi - lo:
                    raise RuntimeError("FusedAdam: the shard of a parameter changed after its state was created")
                if g.dtype not in (torch.float32, torch.float16) or g.numel() != hi - lo:
                    raise RuntimeError("FusedAdam: gradient (slice) must be fp32 or fp16 and match the parameter (slice)")
                work.append((group, p, pv, g, st, lo, hi, mul))
        if not work:
            return None
        # device-side counters: they advance only when the step is not skipped (as torch's capturable Adam does); the increment is
        # formed once per device, not once per parameter
        incs = {}
        for w in work:
            step = w[4]["step"]
            if found_inf is None:
                step += 1.0
                continue
            if step.device not in incs:
                incs[step.device] = 1.0 - found_inf.to(step.device).reshape(())
            step += incs[step.device]
        for group, p, pv, g, st, lo, hi, mul in work:
            b1, b2 = group["betas"]
            shadow = getattr(p, "_enerf_half", None)
            if shadow is not None and (shadow[0].shape != p.shape or shadow[0].device != p.device):
                shadow = None
            sh_ptr = ptr(shadow[0].view(-1)[lo:hi]) if shadow is not None else None
            _lib.call("enerf_adam_step", ptr(pv), ptr(g), _lib.dtype_code(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), hi - lo, ptr(st["step"]),
                      float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                      ptr(grad_scale), ptr(found_inf), float(mul), sh_ptr, stream())
        return None

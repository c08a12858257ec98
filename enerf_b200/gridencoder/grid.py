"""Drop-in for the reference's `gridencoder` package (gridencoder/grid.py).

`GridEncoder` has the reference's constructor, attributes (`embeddings`, `offsets`,
`output_dim`, ...), parameter initialisation and level-offset rule (grid.py:91-135), so state
dicts are interchangeable.  `grid_encode(inputs, embeddings, offsets, per_level_scale,
base_resolution, calc_grad_inputs, gridtype)` is the same autograd function (grid.py:19-88).

What changed underneath (B200-first):
  * the kernel writes / reads the [B, L*C] layout directly: the reference's permute copy after
    forward (grid.py:52) and permute+contiguous before backward (grid.py:70) are gone;
  * under autocast the fp16 copy of the table is cached per parameter version instead of being
    re-cast (25-50 MB) on every call (grid.py:38-39);
  * embedding gradients are accumulated in fp32 (`red.global.add.v2.f32`) and returned in the
    parameter's dtype; the reference accumulates in fp16 atomics when the table is fp16
    (gridencoder.cu:296-302) and lets autograd cast.  Set `ENERF_GRID_GRAD_FP16=1` to reproduce
    the reference's fp16 accumulation.
"""
import os

import numpy as np
import torch
import torch.nn as nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .backend import _backend

_gridtype_to_id = {'hash': 0, 'tiled': 1}

_half_cache = {}   # id(param storage) -> (version, half copy)


def _half_table(emb):
    key = (emb.data_ptr(), emb.numel(), emb.device)
    ver = emb._version
    hit = _half_cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    h = emb.detach().to(torch.half)
    _half_cache.clear()  # a handful of encoders at most; never let stale tables pile up
    _half_cache[key] = (ver, h)
    return h


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0):
        # inputs [B,D] fp32 in [0,1]; embeddings [sum_entries, C]; offsets [L+1] int32 -> [B, L*C]
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = np.log2(per_level_scale)
        H = base_resolution

        table = embeddings
        if torch.is_autocast_enabled('cuda') and C % 2 == 0:
            table = _half_table(embeddings) if embeddings.dtype != torch.half else embeddings
        table = table.contiguous()

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=table.dtype)
        if calc_grad_inputs:
            dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=table.dtype)
        else:
            dy_dx = torch.empty(1, device=inputs.device, dtype=table.dtype)

        _backend.grid_encode_forward(inputs, table, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx, gridtype, 1)

        ctx.save_for_backward(inputs, table, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype]
        ctx.calc_grad_inputs = calc_grad_inputs
        ctx.param_dtype = embeddings.dtype
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, table, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype = ctx.dims
        calc_grad_inputs = ctx.calc_grad_inputs

        grad = grad.contiguous()   # [B, L*C], consumed in place
        if grad.dtype != table.dtype:
            grad = grad.to(table.dtype)

        fp16_accum = os.environ.get('ENERF_GRID_GRAD_FP16', '0') == '1'
        acc_dtype = table.dtype if fp16_accum else torch.float32
        grad_embeddings = torch.zeros(table.shape, dtype=acc_dtype, device=table.device)
        if calc_grad_inputs:
            grad_inputs = torch.zeros_like(inputs, dtype=table.dtype)
        else:
            grad_inputs = torch.zeros(1, device=inputs.device, dtype=table.dtype)

        _backend.grid_encode_backward(grad, inputs, table, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                                      grad_inputs, gridtype, 1)

        if grad_embeddings.dtype != ctx.param_dtype:
            grad_embeddings = grad_embeddings.to(ctx.param_dtype)
        if calc_grad_inputs:
            return grad_inputs.to(inputs.dtype), grad_embeddings, None, None, None, None, None
        return None, grad_embeddings, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash'):
        super().__init__()
        # finest resolution, if given, overrides per_level_scale (grid.py:96-97)
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]

        if level_dim % 2 != 0:
            print('[WARN] detected HashGrid level_dim % 2 != 0, which will cause very slow backward is also enabled fp16! (maybe fix later)')

        # level table sizes: min(2^log2_hashmap_size, (res+1)^D) rounded up to 8 (grid.py:113-123)
        self.max_params = 2 ** log2_hashmap_size
        sizes = []
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            n = min(self.max_params, (resolution + 1) ** input_dim)
            sizes.append(int(np.ceil(n / 8) * 8))
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.n_params = self.offsets[-1] * level_dim

        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype}")

    def forward(self, inputs, bound=1):
        # inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id)
        return outputs.view(prefix_shape + [self.output_dim])

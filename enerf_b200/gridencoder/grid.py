"""Host side of the multiresolution hash-grid encoder — the counterpart of the reference's gridencoder/grid.py with its
public names: `GridEncoder(...)` (attributes `embeddings`, `offsets`, `output_dim`, `per_level_scale`, ...) and
`grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs, gridtype)`.

Kept from the reference so state dicts are interchangeable: the level-size rule (entries of level i =
min(2^log2_hashmap_size, (ceil(base * scale^i) + 1)^D) rounded up to a multiple of 8, grid.py:113-123), the U(-1e-4, 1e-4)
initialisation (:133-135), `desired_resolution` overriding `per_level_scale` (:96-97), inputs mapped from [-bound, bound] to
[0, 1] (:144), half table under autocast when the feature width is even (:38-39).

Different underneath (B200-first):
  * the kernels read and write the [B, L*C] row layout directly, so the permute copy after the forward (grid.py:52) and the
    permute + contiguous before the backward (:70) do not exist;
  * the fp16 copy of the table lives next to the parameter (`param._enerf_half`) instead of being re-cast (25-50 MB) on every
    call: `enerf_b200.optim.FusedAdam` rewrites it inside the Adam kernel (its raw-pointer update does not touch the
    parameter's version counter), every other writer (torch optimizers, `copy_`, `load_state_dict`, EMA swaps) bumps the
    version and triggers a re-cast on the next forward;
  * embedding gradients are accumulated in fp32 and handed back in the parameter's dtype — the reference accumulates with
    fp16 atomics when the table is fp16 (gridencoder.cu:296-302); `set_grad_accumulation('fp16')` reproduces that (parity experiments).
"""
import numpy as np
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from .backend import _backend

_gridtype_to_id = {'hash': 0, 'tiled': 1}
_ROW_LAYOUT = 1                      # out_layout of the C ABI: [B, L*C]
_grad_accumulation = "fp32"


def set_grad_accumulation(mode):
    """'fp32' (default): embedding gradients are accumulated in fp32 reductions; 'fp16': in the table's dtype with fp16 atomics, as
    the reference does under autocast (gridencoder.cu:296-302) — lossy, kept to reproduce the reference's numerics in parity runs."""
    global _grad_accumulation
    if mode not in ("fp32", "fp16"):
        raise ValueError("grad accumulation must be 'fp32' or 'fp16'")
    _grad_accumulation = mode



def half_shadow(param, create=False):
    """The fp16 shadow slot `[tensor, version it mirrors]` of an fp32 table parameter, or None.  One slot per Parameter object
    (two encoders never evict each other; a freed-and-reallocated parameter cannot alias it)."""
    slot = getattr(param, "_enerf_half", None)
    if slot is not None and (slot[0].shape != param.shape or slot[0].device != param.device):
        slot = None                                        # the parameter moved (model.to(...)) or was resized
    if slot is None and create:
        slot = [torch.empty(param.shape, dtype=torch.half, device=param.device), -1]
        param._enerf_half = slot
    return slot


def _half_table(param):
    """fp16 view of the table for the kernels.  Valid while the parameter's version counter is the one the shadow was cast at;
    `FusedAdam` keeps it current without bumping the version (its kernel writes parameter and shadow in one pass)."""
    slot = half_shadow(param, create=True)
    wait = getattr(param, "_enerf_wait", None)
    if wait is not None:
        wait(param)                                        # data-parallel: slices of the table may still be in flight (parallel.ShardedExchange)
    if slot[1] != param._version:
        slot[0].copy_(param.detach())                      # in place: pointers captured in a CUDA graph stay valid
        slot[1] = param._version
    return slot[0]


class _grid_encode(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0, in_add=0.0, in_mul=0.0):
        """inputs [B, D] fp32 in [0, 1] — or, with in_mul != 0, raw positions that the kernels map with x = (raw + in_add) * in_mul —,
        embeddings [entries, C], offsets [L+1] int32  ->  [B, L*C]"""
        x = inputs.contiguous()
        n, dim = x.shape
        levels, width = offsets.shape[0] - 1, embeddings.shape[1]
        log2_scale = np.log2(per_level_scale)
        table = embeddings
        if embeddings.dtype != torch.half and width % 2 == 0 and torch.is_autocast_enabled('cuda'):
            table = _half_table(embeddings)
        table = table.contiguous()
        feats = table.new_empty(n, levels * width)
        jac = table.new_empty(n, levels * dim * width) if calc_grad_inputs else table.new_empty(1)
        geometry = (n, dim, width, levels, log2_scale, base_resolution)
        _backend.grid_encode_forward(x, table, offsets, feats, *geometry, calc_grad_inputs, jac, gridtype, _ROW_LAYOUT, in_add, in_mul)
        ctx.geometry, ctx.gridtype, ctx.want_dx, ctx.param_dtype = geometry, gridtype, calc_grad_inputs, embeddings.dtype
        ctx.xf = (in_add, in_mul)
        ctx.save_for_backward(x, table, offsets, jac)
        return feats

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, d_feats):
        x, table, offsets, jac = ctx.saved_tensors
        d_feats = d_feats.contiguous().to(table.dtype)                 # [B, L*C], read in place by the scatter kernel
        acc_dtype = table.dtype if _grad_accumulation == 'fp16' else torch.float32
        d_table = torch.zeros(table.shape, dtype=acc_dtype, device=table.device)
        d_x = torch.zeros_like(x, dtype=table.dtype) if ctx.want_dx else table.new_empty(1)
        _backend.grid_encode_backward(d_feats, x, table, offsets, d_table, *ctx.geometry, ctx.want_dx, jac, d_x, ctx.gridtype, _ROW_LAYOUT, *ctx.xf)
        return (d_x.to(x.dtype) if ctx.want_dx else None, d_table.to(ctx.param_dtype)) + (None,) * 7


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash'):
        super().__init__()
        if desired_resolution is not None:                              # the finest resolution fixes the growth factor
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        if level_dim % 2 != 0:
            print('[WARN] detected HashGrid level_dim % 2 != 0, which will cause very slow backward is also enabled fp16! (maybe fix later)')
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution, self.log2_hashmap_size = per_level_scale, base_resolution, log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id = gridtype, _gridtype_to_id[gridtype]
        self.max_params = 2 ** log2_hashmap_size

        def entries(level):
            res = int(np.ceil(base_resolution * per_level_scale ** level))
            return int(np.ceil(min(self.max_params, (res + 1) ** input_dim) / 8) * 8)

        starts = np.cumsum([0] + [entries(level) for level in range(num_levels)]).astype(np.int32)
        self.register_buffer('offsets', torch.from_numpy(starts))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(starts[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            self.embeddings.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {finest} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype}")

    def forward(self, inputs, bound=1):
        """inputs [..., input_dim] in [-bound, bound]  ->  [..., num_levels * level_dim]"""
        lead = inputs.shape[:-1]
        if inputs.is_cuda and inputs.dtype == torch.float32 and not inputs.requires_grad and isinstance(bound, (int, float)):
            # the kernels apply (inputs + bound) / (2 * bound) themselves, in ATen's arithmetic: x = (raw + bound) * fp32(1 / fp32(2 bound))
            in_mul = float(np.float32(1.0) / np.float32(2 * bound))
            feats = grid_encode(inputs.reshape(-1, self.input_dim), self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                                False, self.gridtype_id, float(bound), in_mul)
            return feats.view(*lead, self.output_dim)
        unit = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
        feats = grid_encode(unit, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, unit.requires_grad,
                            self.gridtype_id)
        return feats.view(*lead, self.output_dim)

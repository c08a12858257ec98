"""Fused E-NeRF field (nerf/network_ff.py:51-73): sigma-net -> trunc_exp -> [SH | geo_feat | 0] ->
colour-net -> sigmoid, with the glue (exp, SH encoder, cat, zeros_like, sigmoid and their
backward) folded into the heads / prologues of the tcgen05 MLP kernels.  Values follow the
unfused module chain step by step (same fp16 rounding points); what disappears is ~1 GB of HBM
traffic and ~20 small kernels per step.

`fused_field(feat, dirs, w_sigma, w_color, nl_sigma, nl_color, n_ch, training)` ->
(sigma [S] fp32, rgb [S,n_ch] fp32).  Differentiable in feat, w_sigma, w_color.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import ptr, stream


def shapes_eligible(hidden_dim, hidden_dim_color, in_dim, in_dim_color, geo_feat_dim, sh_degree, n_ch, activation):
    return (hidden_dim == 64 and hidden_dim_color == 64 and in_dim == 32 and in_dim_color == 32 and geo_feat_dim == 15 and sh_degree == 4
            and 1 <= n_ch <= 4 and activation == 0)


def eligible(feat, dirs, *shape_args):
    return (feat.is_cuda and feat.dtype == torch.float16 and feat.shape[0] % 128 == 0 and feat.shape[0] > 0
            and dirs.shape[0] == feat.shape[0] and shapes_eligible(*shape_args))


# True (default): the training forward stores nothing but its inputs and outputs; the backward kernels recompute the hidden
# activations of every 128-sample tile on the tensor cores (bit-identical values).  False: store / reload forward_buffer
# (2.2 GB written + 2.8 GB read per step on the bench workload, which made the MLP kernels HBM-bound).
RECOMPUTE = True


class _FusedField(Function):
    @staticmethod
    def forward(ctx, feat, dirs, w_sigma, w_color, nl_sigma, nl_color, n_ch, training):
        S = feat.shape[0]
        dev = feat.device
        feat = feat.contiguous()
        dirs = dirs.contiguous().float()
        ws, wc = w_sigma.detach().half().contiguous(), w_color.detach().half().contiguous()
        sigma = torch.empty(S, dtype=torch.float32, device=dev)
        cin = torch.empty(S, 32, dtype=torch.float16, device=dev)
        rgb = torch.empty(S, n_ch, dtype=torch.float32, device=dev)
        store = training and not RECOMPUTE
        fb_s = torch.empty(nl_sigma, S, 64, dtype=torch.float16, device=dev) if store else None
        fb_c = torch.empty(nl_color, S, 64, dtype=torch.float16, device=dev) if store else None
        _lib.call("enerf_field_sigma_forward", ptr(feat), ptr(ws), ptr(dirs), S, nl_sigma, ptr(fb_s), ptr(sigma), ptr(cin), stream())
        _lib.call("enerf_field_color_forward", ptr(cin), ptr(wc), S, nl_color, n_ch, ptr(fb_c), ptr(rgb), None, stream())
        if training:
            saved = [feat, ws, wc, sigma, cin, rgb] + ([fb_s, fb_c] if store else [])
            ctx.save_for_backward(*saved)
            ctx.meta = (nl_sigma, nl_color, n_ch, w_sigma.dtype, w_color.dtype)
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        saved = ctx.saved_tensors
        feat, ws, wc, sigma, cin, rgb = saved[:6]
        fb_s, fb_c = (saved[6], saved[7]) if len(saved) == 8 else (None, None)
        nl_sigma, nl_color, n_ch, dt_s, dt_c = ctx.meta
        S = feat.shape[0]
        dev = feat.device
        g_sigma = torch.zeros_like(sigma) if g_sigma is None else g_sigma.contiguous().float()
        g_rgb = torch.zeros_like(rgb) if g_rgb is None else g_rgb.contiguous().float()
        dcin = torch.empty(S, 32, dtype=torch.float16, device=dev)
        dfeat = torch.empty(S, 32, dtype=torch.float16, device=dev)
        gw_c = torch.empty(wc.numel(), dtype=torch.float32, device=dev)
        gw_s = torch.empty(ws.numel(), dtype=torch.float32, device=dev)
        _lib.call("enerf_field_color_backward", ptr(g_rgb), ptr(rgb), n_ch, ptr(cin), ptr(wc), ptr(fb_c), S, nl_color, ptr(dcin), ptr(gw_c), None, stream())
        _lib.call("enerf_field_sigma_backward", ptr(g_sigma), ptr(sigma), ptr(dcin), ptr(feat), ptr(ws), ptr(fb_s), S, nl_sigma, ptr(dfeat),
                  ptr(gw_s), stream())
        return dfeat, None, gw_s.to(dt_s), gw_c.to(dt_c), None, None, None, None


def fused_field(feat, dirs, w_sigma, w_color, nl_sigma, nl_color, n_ch, training):
    return _FusedField.apply(feat, dirs, w_sigma, w_color, nl_sigma, nl_color, n_ch, training)


# --------------------------------------------------------------------------------------------
# Inference (torch.no_grad): encoder + both nets as ONE kernel (csrc/field_infer.cu).  The [S,32] feature rows and the [S,32]
# colour-net input rows never reach HBM; sigma and rgb are the bits `encoder -> fused_field` produces.
def infer_eligible(x, dirs, encoder, *shape_args):
    """raw positions [S,3] fp32, directions [S,3], a hash / tiled GridEncoder of 16 levels x 2 features over 3-D inputs, FFMLP 2 + 3 layers"""
    from .gridencoder.grid import GridEncoder
    return (isinstance(encoder, GridEncoder) and encoder.input_dim == 3 and encoder.num_levels == 16 and encoder.level_dim == 2
            and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == 3 and dirs.shape == x.shape
            and shapes_eligible(*shape_args))


@torch.no_grad()
def fused_infer(x, dirs, encoder, bound, w_sigma, w_color, n_ch, alive=None, out=None):
    """x [S,3] in [-bound, bound], dirs [S,3] -> (sigma [S] fp32, rgb [S,n_ch] fp32); no autograd graph.
    `alive` = (int32 device scalar u, rows_per_unit): only the first min(S, u * rows_per_unit) rows are evaluated and written (the
    inference loop's alive rays x steps of the round); the other rows of the results are left as they are (uninitialised, or what
    `out` = (sigma, rgb) — contiguous fp32 tensors to write into — held)."""
    import numpy as np
    from .gridencoder.grid import _half_table
    S = x.shape[0]
    dev = x.device
    x = x.contiguous()
    dirs = dirs.contiguous().float()
    table = encoder.embeddings if encoder.embeddings.dtype == torch.float16 else _half_table(encoder.embeddings)
    ws, wc = w_sigma.detach().half().contiguous(), w_color.detach().half().contiguous()
    if out is not None:
        sigma, rgb = out
        if not (sigma.shape == (S,) and rgb.shape == (S, n_ch) and sigma.dtype == rgb.dtype == torch.float32 and sigma.is_contiguous()
                and rgb.is_contiguous() and sigma.device == rgb.device == dev):
            raise ValueError("fused_infer: out = (sigma [S], rgb [S, n_ch]), contiguous fp32 on the inputs' device")
    else:
        sigma = torch.empty(S, dtype=torch.float32, device=dev)
        rgb = torch.empty(S, n_ch, dtype=torch.float32, device=dev)
    in_mul = float(np.float32(1.0) / np.float32(2 * bound))          # GridEncoder.forward's (x + bound) / (2 bound) in ATen's arithmetic
    count, per_unit = alive if alive is not None else (None, 0)
    if count is not None and not (count.is_cuda and count.dtype == torch.int32 and count.numel() == 1 and int(per_unit) > 0):
        raise ValueError("fused_infer: alive = (int32 CUDA scalar, rows per unit > 0)")
    _lib.call("enerf_field_infer_alive", ptr(x), float(bound), in_mul, ptr(dirs), ptr(table), ptr(encoder.offsets), encoder.num_levels, encoder.level_dim,
              float(np.log2(encoder.per_level_scale)), int(encoder.base_resolution), int(encoder.gridtype_id), ptr(ws), 2, ptr(wc), 3, S, n_ch,
              ptr(sigma), ptr(rgb), ptr(count), int(per_unit), stream())
    return sigma, rgb


# --------------------------------------------------------------------------------------------
# The torch-topology field of nerf/network.py:104-199 — what every shipped E-NeRF config runs (ff = False): sigma-net
# Linear(32,64)-ReLU-Linear(64,16), colour-net Linear(31,64)-ReLU-Linear(64,64)-ReLU-Linear(64,C) on the `weights > 1e-4` samples.
# Same tcgen05 kernels as the FFMLP path (one hidden-to-hidden matmul fewer per net); the weights are the nn.Linear matrices
# concatenated in FFMLP order, the colour-net's first matrix padded with a zero 32nd column (its input row is [SH | geo_feat | 0])
# and its last one with zero rows up to 16.
def torch_topology_eligible(hidden_dim, num_layers, num_layers_color, in_dim, in_dim_dir, geo_feat_dim, sh_degree, n_ch):
    return (hidden_dim == 64 and num_layers == 2 and num_layers_color == 3 and in_dim == 32 and in_dim_dir == 16 and geo_feat_dim == 15
            and sh_degree == 4 and 1 <= n_ch <= 4)


def flat_sigma_weights(layers):
    """[Linear(32,64), Linear(64,16)] -> flat [64*32 + 16*64]"""
    return torch.cat([l.weight.reshape(-1) for l in layers])


def flat_color_weights(layers):
    """[Linear(31,64), Linear(64,64), Linear(64,C)] -> flat [64*32 + 64*64 + 16*64] (zero column / zero rows added)"""
    w0, w1, w2 = (l.weight for l in layers)
    w0 = torch.nn.functional.pad(w0, (0, 32 - w0.shape[1]))
    w2 = torch.nn.functional.pad(w2, (0, 0, 0, 16 - w2.shape[0]))
    return torch.cat([w0.reshape(-1), w1.reshape(-1), w2.reshape(-1)])


class _Density(Function):
    """feat [B,32] fp16, flat sigma-net weights -> sigma [B] fp32 (= trunc_exp(h[0])), h [B,16] fp16 (geo_feat = h[:,1:])."""

    @staticmethod
    def forward(ctx, feat, w_sigma, num_layers):
        B = feat.shape[0]
        feat = feat.contiguous()
        ws = w_sigma.detach().half().contiguous()
        sigma = torch.empty(B, dtype=torch.float32, device=feat.device)
        h = torch.empty(B, 16, dtype=torch.float16, device=feat.device)
        _lib.call("enerf_field_density_forward", ptr(feat), ptr(ws), B, num_layers, ptr(sigma), ptr(h), stream())
        ctx.save_for_backward(feat, ws, sigma)
        ctx.meta = (num_layers, w_sigma.dtype)
        return sigma, h

    @staticmethod
    def backward(ctx, g_sigma, g_h):
        feat, ws, sigma = ctx.saved_tensors
        num_layers, dt = ctx.meta
        B = feat.shape[0]
        g_sigma = None if g_sigma is None else g_sigma.contiguous().float()
        g_h = None if g_h is None else g_h.contiguous().half()
        dfeat = torch.empty_like(feat)
        gw = torch.empty(ws.numel(), dtype=torch.float32, device=feat.device)
        _lib.call("enerf_field_density_backward", ptr(g_sigma), ptr(sigma), ptr(g_h), ptr(feat), ptr(ws), B, num_layers, ptr(dfeat), ptr(gw), stream())
        return dfeat, gw.to(dt), None


def density_head(feat, w_sigma, num_layers=1):
    return _Density.apply(feat, w_sigma, num_layers)


def density_only(feat, w_sigma, num_layers):
    """sigma [B] fp32 alone (no gradient): the occupancy-grid refresh"""
    B = feat.shape[0]
    sigma = torch.empty(B, dtype=torch.float32, device=feat.device)
    _lib.call("enerf_field_density_forward", ptr(feat.contiguous()), ptr(w_sigma.detach().half().contiguous()), B, num_layers, ptr(sigma), None, stream())
    return sigma


class _MaskedColor(Function):
    """h [B,16] fp16, dirs [B/dir_div,3] fp32, idx [cap] int32 (selected samples, increasing; the first `count` entries are valid), flat
    colour-net weights -> rgbs [B,n_ch] fp32 with sigmoid(colour-net) on the selected rows and zeros elsewhere (network.py:171-199).
    `count`: int32 device scalar (from `raymarching.compact_mask`) or None (= all of idx).  The count never travels to the host: every
    kernel reads it from the device and skips the tiles / rows beyond it, so the whole render stays capturable in a CUDA graph."""

    @staticmethod
    def forward(ctx, h, dirs, dir_div, idx, count, w_color, n_ch, sh_scale):
        B, cap = h.shape[0], idx.shape[0]
        dev = h.device
        cap_pad = -(-cap // 128) * 128
        h = h.contiguous()
        wc = w_color.detach().half().contiguous()
        cin = torch.empty(cap_pad, 32, dtype=torch.float16, device=dev)
        _lib.call("enerf_field_color_inputs", ptr(dirs), dir_div, ptr(h), ptr(idx), cap, cap_pad, float(sh_scale), ptr(cin), ptr(count), stream())
        rgb_c = torch.empty(cap_pad, n_ch, dtype=torch.float32, device=dev)
        _lib.call("enerf_field_color_forward", ptr(cin), ptr(wc), cap_pad, 2, n_ch, None, ptr(rgb_c), ptr(count), stream())
        rgbs = torch.zeros(B, n_ch, dtype=torch.float32, device=dev)
        _lib.call("enerf_scatter_rows", ptr(rgb_c), ptr(idx), cap, 4 * n_ch, ptr(rgbs), ptr(count), stream())
        ctx.save_for_backward(cin, wc, rgb_c, idx, count if count is not None else idx.new_empty(0))
        ctx.meta = (B, cap, cap_pad, n_ch, w_color.dtype, count is not None)
        return rgbs

    @staticmethod
    def backward(ctx, g_rgbs):
        cin, wc, rgb_c, idx, count = ctx.saved_tensors
        B, cap, cap_pad, n_ch, dt, has_count = ctx.meta
        count = count if has_count else None
        dev = cin.device
        g_rgbs = g_rgbs.contiguous().float()
        g_c = torch.empty(cap_pad, n_ch, dtype=torch.float32, device=dev)
        _lib.call("enerf_gather_rows", ptr(g_rgbs), ptr(idx), cap, cap_pad, 4 * n_ch, ptr(g_c), ptr(count), stream())
        dcin = torch.empty(cap_pad, 32, dtype=torch.float16, device=dev)
        gw = torch.empty(wc.numel(), dtype=torch.float32, device=dev)
        _lib.call("enerf_field_color_backward", ptr(g_c), ptr(rgb_c), n_ch, ptr(cin), ptr(wc), None, cap_pad, 2, ptr(dcin), ptr(gw), ptr(count), stream())
        g_h = torch.zeros(B, 16, dtype=torch.float16, device=dev)
        _lib.call("enerf_field_color_inputs_backward", ptr(dcin), ptr(idx), cap, ptr(g_h), ptr(count), stream())
        return g_h, None, None, None, None, gw.to(dt), None, None


def masked_color(h, dirs, dir_div, idx, count, w_color, n_ch, sh_scale=1.0):
    return _MaskedColor.apply(h, dirs.contiguous().float(), int(dir_div), idx, count, w_color, int(n_ch), float(sh_scale))

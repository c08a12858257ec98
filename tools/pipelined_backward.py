"""EXPERIMENT (measured slower, not part of the product): encoder + field as ONE autograd node whose backward overlaps the
hash-grid scatter of sample chunk k (second stream) with the MLP backward of chunk k+1.  Moved here from enerf_b200/field.py in
round 2; used by tools/overlap_probe.py only.  Result: profiles/r1_30_overlap_probe_two_streams.json."""
import os

import numpy as np
import torch
from torch.autograd import Function

from enerf_b200 import _lib
from enerf_b200._lib import ptr, stream

# --------------------------------------------------------------------------------------------
# Encoder + field as ONE autograd node with a pipelined backward.
#
# The two halves of the backward have different bottlenecks: the MLP backward kernels are bound by the SM's shared-memory /
# tensor pipes (persistent, one CTA per SM), the hash-grid scatter by L2 reduction throughput (it needs few SMs).  Run back to
# back they cost their sum.  Here the samples are cut into chunks; the MLP backward of chunk k+1 runs on the current stream on
# a capped number of SMs while the scatter of chunk k runs on a second stream on the rest.  Same kernels, same sums (fp32
# reductions in a different order); CUDA-graph capturable (fork / join through events).
#
# Measured on B200 (profiles/r1_30_overlap_probe_two_streams.json, 3.29 M samples): SLOWER than back to back in every setting
# (1.32 ms -> 1.46 / 1.66 / 1.84 ms with 2 / 4 / 8 chunks; capping the MLP grid makes it worse) — both kernels need all SMs, and an
# MLP CTA cannot be placed until the scatter CTAs on its SM have drained.  Kept as a tested experiment: ENERF_PIPELINE=1 turns it
# on (default off); ENERF_PIPELINE_CHUNKS / _MLP_CTAS / _SCATTER_BLOCK.
PIPELINE = os.environ.get("ENERF_PIPELINE", "0") == "1"
PIPELINE_CHUNKS = int(os.environ.get("ENERF_PIPELINE_CHUNKS", "4"))
PIPELINE_MLP_CTAS = int(os.environ.get("ENERF_PIPELINE_MLP_CTAS", "120"))
PIPELINE_SCATTER_BLOCK = int(os.environ.get("ENERF_PIPELINE_SCATTER_BLOCK", "0"))
_side_streams = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key)
    return _side_streams[key]


def encoder_eligible(encoder, x):
    """hash / tiled grid with D = 3, 16 levels x 2 features (the 32-wide sigma-net input), no input gradients"""
    return (type(encoder).__name__ == "GridEncoder" and encoder.input_dim == 3 and encoder.level_dim == 2 and encoder.num_levels == 16
            and x.is_cuda and x.shape[-1] == 3 and not x.requires_grad and x.numel() // 3 % 128 == 0 and x.numel() > 0)


def pipelined_backward(g_sigma, g_rgb, sigma, rgb, cin, feat, x, table, offsets, geometry, gridtype, ws, wc, nl_sigma, nl_color, n_ch,
                       acc_dtype=torch.float32, chunks=None, mlp_ctas=None, scatter_block=None):
    """colour-net bwd -> sigma-net bwd -> hash-grid scatter over `chunks` sample ranges, the scatter of a range on a second stream
    while the MLP kernels work on the next one.  Returns (d_table [entries, C] acc_dtype, gw_sigma fp32, gw_color fp32)."""
    from enerf_b200.backends import gridencoder_backend as GB
    chunks = PIPELINE_CHUNKS if chunks is None else chunks
    mlp_ctas = PIPELINE_MLP_CTAS if mlp_ctas is None else mlp_ctas
    scatter_block = PIPELINE_SCATTER_BLOCK if scatter_block is None else scatter_block
    S, dev = feat.shape[0], feat.device
    _, dim, width, levels, log2_scale, base_resolution = geometry
    step = -(-(S // 128) // max(1, chunks)) * 128
    n_used = -(-S // step)
    main, side = torch.cuda.current_stream(dev), _side_stream(dev)
    d_table = torch.zeros(table.shape, dtype=acc_dtype, device=dev)
    dcin = torch.empty(S, 32, dtype=torch.float16, device=dev)
    dfeat = torch.empty(S, 32, dtype=torch.float16, device=dev)
    gw_c = torch.empty(n_used, wc.numel(), dtype=torch.float32, device=dev)
    gw_s = torch.empty(n_used, ws.numel(), dtype=torch.float32, device=dev)
    dummy = table.new_zeros(1)
    fork = torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    _lib.call("enerf_ffmlp_set_max_ctas", mlp_ctas)
    _lib.call("enerf_grid_set_backward_block", scatter_block)
    try:
        for k in range(n_used):
            lo, hi = k * step, min(S, (k + 1) * step)
            n = hi - lo
            _lib.call("enerf_field_color_backward", ptr(g_rgb[lo:hi]), ptr(rgb[lo:hi]), n_ch, ptr(cin[lo:hi]), ptr(wc), None, n, nl_color,
                      ptr(dcin[lo:hi]), ptr(gw_c[k]), None, stream())
            _lib.call("enerf_field_sigma_backward", ptr(g_sigma[lo:hi]), ptr(sigma[lo:hi]), ptr(dcin[lo:hi]), ptr(feat[lo:hi]), ptr(ws), None, n,
                      nl_sigma, ptr(dfeat[lo:hi]), ptr(gw_s[k]), stream())
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                GB.grid_encode_backward(dfeat[lo:hi], x[lo:hi], table, offsets, d_table, n, dim, width, levels, log2_scale, base_resolution,
                                        False, dummy, dummy, gridtype, 1)
    finally:
        _lib.call("enerf_ffmlp_set_max_ctas", 0)
        _lib.call("enerf_grid_set_backward_block", 0)
    join = torch.cuda.Event()
    join.record(side)
    main.wait_event(join)
    return d_table, gw_s.sum(0), gw_c.sum(0)


class _EncodedField(Function):
    """hash-grid gather + fused field (forward identical to `grid_encode` followed by `fused_field`); backward pipelined."""

    @staticmethod
    def forward(ctx, x, dirs, embeddings, offsets, per_level_scale, base_resolution, gridtype, w_sigma, w_color, nl_sigma, nl_color, n_ch,
                training):
        from enerf_b200.backends import gridencoder_backend as GB
        from enerf_b200.gridencoder.grid import _half_table
        x = x.contiguous().float()
        S, dev = x.shape[0], x.device
        table = (embeddings.detach() if embeddings.dtype == torch.half else _half_table(embeddings)).contiguous()
        levels, width = offsets.shape[0] - 1, embeddings.shape[1]
        geometry = (S, 3, width, levels, np.log2(per_level_scale), base_resolution)
        feat = table.new_empty(S, levels * width)
        GB.grid_encode_forward(x, table, offsets, feat, *geometry, False, table.new_empty(1), gridtype, 1)
        dirs = dirs.contiguous().float()
        ws, wc = w_sigma.detach().half().contiguous(), w_color.detach().half().contiguous()
        sigma = torch.empty(S, dtype=torch.float32, device=dev)
        cin = torch.empty(S, 32, dtype=torch.float16, device=dev)
        rgb = torch.empty(S, n_ch, dtype=torch.float32, device=dev)
        _lib.call("enerf_field_sigma_forward", ptr(feat), ptr(ws), ptr(dirs), S, nl_sigma, None, ptr(sigma), ptr(cin), stream())
        _lib.call("enerf_field_color_forward", ptr(cin), ptr(wc), S, nl_color, n_ch, None, ptr(rgb), None, stream())
        if training:
            ctx.save_for_backward(x, table, offsets, feat, ws, wc, sigma, cin, rgb)
            ctx.meta = (geometry, gridtype, nl_sigma, nl_color, n_ch, embeddings.dtype, w_sigma.dtype, w_color.dtype)
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        x, table, offsets, feat, ws, wc, sigma, cin, rgb = ctx.saved_tensors
        geometry, gridtype, nl_sigma, nl_color, n_ch, dt_e, dt_s, dt_c = ctx.meta
        g_sigma = torch.zeros_like(sigma) if g_sigma is None else g_sigma.contiguous().float()
        g_rgb = torch.zeros_like(rgb) if g_rgb is None else g_rgb.contiguous().float()
        acc_dtype = torch.float32
        d_table, gw_s, gw_c = pipelined_backward(g_sigma, g_rgb, sigma, rgb, cin, feat, x, table, offsets, geometry, gridtype, ws, wc,
                                                 nl_sigma, nl_color, n_ch, acc_dtype)
        return (None, None, d_table.to(dt_e), None, None, None, None, gw_s.to(dt_s), gw_c.to(dt_c), None, None, None, None)


def encoded_field(x_unit, dirs, encoder, w_sigma, w_color, nl_sigma, nl_color, n_ch, training):
    """x_unit [S,3] in [0,1] -> (sigma [S] fp32, rgb [S,n_ch] fp32); differentiable in the table and both weight vectors."""
    return _EncodedField.apply(x_unit, dirs, encoder.embeddings, encoder.offsets, encoder.per_level_scale, encoder.base_resolution,
                               encoder.gridtype_id, w_sigma, w_color, nl_sigma, nl_color, n_ch, training)

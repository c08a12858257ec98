#!/usr/bin/env python
"""Does pinning the hash table (gather) / the gradient table (scatter) in L2 with an access-policy window help?  Times both kernels on
the bench workload's marched samples with and without a persisting window on the stream (L2 flushed between launches either way, so
"without" is the cold-table case and "with" the best case; in the training step the table is usually still L2-resident from Adam)."""
import json
import os
import sys

import numpy as np
import torch
from cuda.bindings import runtime as cudart

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import synthetic  # noqa: E402
from enerf_b200 import raymarching as rm  # noqa: E402
from enerf_b200.backends import gridencoder_backend as GB  # noqa: E402
from enerf_b200.gridencoder import GridEncoder  # noqa: E402


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def window(stream, tensor, on):
    attr = cudart.cudaStreamAttrValue()
    attr.accessPolicyWindow.base_ptr = tensor.data_ptr()
    attr.accessPolicyWindow.num_bytes = tensor.numel() * tensor.element_size() if on else 0
    attr.accessPolicyWindow.hitRatio = 1.0
    attr.accessPolicyWindow.hitProp = cudart.cudaAccessProperty.cudaAccessPropertyPersisting
    attr.accessPolicyWindow.missProp = cudart.cudaAccessProperty.cudaAccessPropertyStreaming
    err, = cudart.cudaStreamSetAttribute(stream.cuda_stream, cudart.cudaStreamAttrID.cudaLaunchAttributeAccessPolicyWindow, attr)
    return int(err)


def main():
    dev = torch.device("cuda", 0)
    bound = 3
    err, = cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitPersistingL2CacheSize, 64 << 20)
    res = {"set_limit_err": int(err)}
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        bits = torch.from_numpy(synthetic.packbits_np(synthetic.ball_density_grid(bound, 3))).to(dev)
        o, d = synthetic.random_rays(4096, bound, seed=100)
        o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
        aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
        nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        xyzs, _, _, _ = rm.march_rays_train(o, d, float(bound), bits, 3, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
        S = xyzs.shape[0] // 128 * 128
        x = ((xyzs[:S] + bound) / (2 * bound)).contiguous()
        enc = GridEncoder(desired_resolution=2048 * bound).to(dev)
        table = (torch.rand_like(enc.embeddings) - 0.5).half().contiguous()
        out = torch.empty(S, 32, dtype=torch.half, device=dev)
        grad = (torch.randn(S, 32, device=dev) * 0.01).half()
        gtab = torch.zeros(enc.embeddings.shape, dtype=torch.float32, device=dev)
        dummy = torch.empty(1, dtype=torch.half, device=dev)
        log2s = float(np.log2(enc.per_level_scale))
        res["samples"] = S
        fwd = lambda: GB.grid_encode_forward(x, table, enc.offsets, out, S, 3, 2, 16, log2s, 16, False, dummy, 0, 1)
        bwd = lambda: GB.grid_encode_backward(grad, x, table, enc.offsets, gtab, S, 3, 2, 16, log2s, 16, False, dummy, dummy, 0, 1)
        for on in (False, True, False, True):
            e1 = window(st, table, on)
            f = timeit(fwd)
            e2 = window(st, gtab, on)
            b = timeit(bwd)
            res.setdefault("runs", []).append({"persisting_window": on, "gather_ms": f, "scatter_ms": b, "err": [e1, e2]})
    print(json.dumps(res))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Walking scatter: time vs threads per CTA (enerf_grid_set_backward_block) on the bench workload's marched samples."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import _lib, synthetic  # noqa: E402
from enerf_b200 import raymarching as rm  # noqa: E402
from enerf_b200.backends import gridencoder_backend as GB  # noqa: E402
from enerf_b200.gridencoder import GridEncoder  # noqa: E402

dev = torch.device("cuda", 0)
bound, cascade, n_rays = 3, 3, 4096
bits = torch.from_numpy(synthetic.packbits_np(synthetic.ball_density_grid(bound, cascade))).to(dev)
o, d = synthetic.random_rays(n_rays, bound, seed=100)
o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
counter = torch.zeros(2, dtype=torch.int32, device=dev)
xyzs, _, _, _ = rm.march_rays_train(o, d, float(bound), bits, cascade, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
S = xyzs.shape[0]
x = ((xyzs + bound) / (2 * bound)).contiguous()
enc = GridEncoder(desired_resolution=2048 * bound).to(dev)
table = enc.embeddings.detach().half().contiguous()
grad = (torch.randn(S, 32, device=dev) * 1e-2).half()
gt = torch.zeros(table.shape, dtype=torch.float32, device=dev)
dummy = torch.empty(1, dtype=torch.half, device=dev)
log2s = float(np.log2(enc.per_level_scale))
res = {"samples": S}
for block in (256, 192, 128, 64, 256):
    _lib.call("enerf_grid_set_backward_block", block)
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gt.zero_()
        a.record()
        GB.grid_encode_backward(grad, x, table, enc.offsets, gt, S, 3, 2, 16, log2s, 16, False, dummy, dummy, 0, 1)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    res.setdefault(f"block_{block}_ms", []).append(round(float(np.median(ts)), 4))
_lib.call("enerf_grid_set_backward_block", 0)
print(json.dumps(res))

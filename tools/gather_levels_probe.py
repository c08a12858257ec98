#!/usr/bin/env python
"""Upper bound on what staging the dense coarse levels of the hash grid in shared memory could buy the gather (VERDICT r1 item 6:
"evaluate smem staging of dense levels 0-1"): time the gather on the bench workload's marched samples with all 16 levels, with the two
coarsest levels removed (14 levels, 35 -> 6144) and with only the 11 hashed levels (77 -> 6144).  The difference is the whole cost of the
removed levels — arithmetic, L1 traffic and all; staging them in shared memory could at best remove their memory part."""
import json
import os
import sys

import numpy as np
import torch

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


def main():
    dev = torch.device("cuda", 0)
    bound = 3
    bits = torch.from_numpy(synthetic.packbits_np(synthetic.ball_density_grid(bound, 3))).to(dev)
    o, d = synthetic.random_rays(4096, bound, seed=100)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, _, _, _ = rm.march_rays_train(o, d, float(bound), bits, 3, 128, nears, fars, counter, -1, True, 128, False, 0, 1024)
    S = xyzs.shape[0] // 128 * 128
    x = ((xyzs[:S] + bound) / (2 * bound)).contiguous()
    res = {"samples": S}
    for name, levels, base in (("all_16_levels", 16, 16), ("without_levels_0_1", 14, 35), ("hashed_levels_only", 11, 117)):
        enc = GridEncoder(num_levels=levels, base_resolution=base, desired_resolution=2048 * bound).to(dev)
        table = (torch.rand_like(enc.embeddings) - 0.5).half().contiguous()
        out = torch.empty(S, levels * 2, dtype=torch.half, device=dev)
        dummy = torch.empty(1, dtype=torch.half, device=dev)
        log2s = float(np.log2(enc.per_level_scale))
        ms = timeit(lambda: GB.grid_encode_forward(x, table, enc.offsets, out, S, 3, 2, levels, log2s, base, False, dummy, 0, 1))
        dense = int(((enc.offsets[1:] - enc.offsets[:-1]) < (1 << 19)).sum())
        res[name] = {"ms": ms, "levels": levels, "dense_levels": dense, "ms_per_level": ms / levels}
    a, b = res["all_16_levels"]["ms"], res["without_levels_0_1"]["ms"]
    res["cost_of_levels_0_1_ms"] = a - b
    res["share_of_gather"] = (a - b) / a
    print(json.dumps(res))


if __name__ == "__main__":
    main()

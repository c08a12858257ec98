#!/usr/bin/env python
"""Full-frame inference (BASELINE configs[3]) vs the per-round sample budget and the host-sync interval of the mirror's inference loop."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    D = bench.Dist()
    torch.manual_seed(0)
    model = bench.make_ff_model(dev, bench.BOUND)
    res = []
    for log2, every in ((0, 1), (23, 1), (23, 4), (24, 4), (25, 4), (26, 4)):
        model.inference_batch_samples = (1 << log2) if log2 else 0
        model.inference_sync_every = every
        r = bench.render_bench(model, dev, D, frames=2)
        res.append({"batch_samples": model.inference_batch_samples, "sync_every": every, "frame_ms": r["frame_ms"], "msamples_per_s": r["msamples_per_s"],
                    "iterations": r["iterations"], "host_syncs": r["host_syncs"], "samples": r["samples_shaded"]})
        print(json.dumps(res[-1]), flush=True)
    print(json.dumps({"render_probe": res}))


if __name__ == "__main__":
    main()

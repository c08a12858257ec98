#!/usr/bin/env python
"""Where the 800x800 inference frame (BASELINE configs[3]) spends its time: CUDA-event pairs around every C-ABI call of one frame
(`_lib.profile_start`), next to the frame's wall time on the device."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from enerf_b200 import _lib  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    D = bench.Dist()
    torch.manual_seed(0)
    model = bench.make_ff_model(dev, bench.BOUND)
    r0 = bench.render_bench(model, dev, D, frames=2)
    _lib.profile_start()
    bench.render_bench(model, dev, D, frames=1)                # two frames with an event pair around every call
    prof = _lib.profile_stop()
    calls = {k: {"calls_per_frame": n / 2, "ms_per_frame": ms / 2} for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    print(json.dumps({"frame_ms": r0["frame_ms"], "iterations": r0["iterations"], "samples": r0["samples_shaded"],
                      "c_abi_ms_per_frame": sum(v["ms_per_frame"] for v in calls.values()), "calls": calls}))


if __name__ == "__main__":
    main()

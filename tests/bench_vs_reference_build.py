#!/usr/bin/env python
"""CLI around oracle/gpu_bar.py: kernel times of the reference's own CUDA build next to this repo's kernels.

  python tests/bench_vs_reference_build.py [--rays 4096] [--iters 5] [--out gpurun_out/x.json]
Every row is printed as soon as it is measured; a failing row does not stop the rest."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gpu_bar  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--bound", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    print(json.dumps(gpu_bar.measure(a.rays, a.iters, a.bound, a.out, verbose=True)))

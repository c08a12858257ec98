#!/usr/bin/env python
"""Cross-check the tcgen05 backward kernels (TMA operand loads vs per-thread loads) on the GPU."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import _lib
from enerf_b200.backends import ffmlp_backend as FB

dev = "cuda"
torch.manual_seed(0)
for B in (128, 256, 128 * 148 * 3 + 128 * 5):
    for nl in (2, 3):
        for positive in (True, False):
            nw = 64 * (32 + 64 * (nl - 1) + 16)
            w = ((torch.rand(nw, device=dev) * 2 - 1) * (3 / 64) ** 0.5)
            x = torch.randn(B, 32, device=dev) * 0.5
            if positive:
                w, x = w.abs(), x.abs()
            w, x = w.half(), x.half()
            g = (torch.randn(B, 16, device=dev) * 0.1).half()
            out = torch.empty(B, 16, device=dev, dtype=torch.half)
            fb = torch.empty(nl, B, 64, device=dev, dtype=torch.half)
            FB.ffmlp_forward(x, w, B, 32, 16, 64, nl, 0, 6, fb, out)
            res = {}
            for path in (2, 0, "rc"):
                _lib.call("enerf_ffmlp_set_path", 0 if path == "rc" else path)
                gi = torch.zeros(B, 32, device=dev, dtype=torch.half)
                gw = torch.zeros(nw, device=dev, dtype=torch.float32)
                FB.ffmlp_backward(g, x, w, None if path == "rc" else fb, B, 32, 16, 64, nl, 0, 6, True, None, gi, gw)
                torch.cuda.synchronize()
                res[path] = (gi.float().cpu(), gw.cpu())
            _lib.call("enerf_ffmlp_set_path", 0)
            gi_ref, gw_ref = res[2]
            gi_new, gw_new = res[0]
            gi_rc, gw_rc = res["rc"]
            print(f"   recompute vs stored (TMA): gi err {float((gi_rc - gi_new).abs().max()):.4g}  gw err {float((gw_rc - gw_new).abs().max()):.4g} / {float(gw_new.abs().max()):.4g}")
            seg = [("W0", 0, 64 * 32)] + [(f"Wh{j}", 64 * 32 + j * 4096, 64 * 32 + (j + 1) * 4096) for j in range(nl - 1)] + [("Wl", nw - 1024, nw)]
            msg = f"B={B} nl={nl} positive={positive}: gi err {float((gi_new - gi_ref).abs().max()):.4g} / {float(gi_ref.abs().max()):.4g}"
            for name, a, b in seg:
                msg += f" | {name} {float((gw_new[a:b] - gw_ref[a:b]).abs().max()):.4g}/{float(gw_ref[a:b].abs().max()):.4g}"
            print(msg, flush=True)
            if B == 128 and nl == 2:
                e = (gi_new - gi_ref).abs()
                print("  per-column max err:", [round(float(v), 3) for v in e.max(0).values])
                print("  per-row-group(8) max err:", [round(float(v), 3) for v in e.view(16, 8, 32).amax((1, 2))])
                print("  new[0,:8]", gi_new[0, :8].tolist(), "\n  ref[0,:8]", gi_ref[0, :8].tolist())
                # is new == ref of some other row / a multiple?
                ratio = (gi_new / (gi_ref + 1e-9))
                print("  ratio[0,:8]", ratio[0, :8].tolist())

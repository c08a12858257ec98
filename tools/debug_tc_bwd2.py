#!/usr/bin/env python
"""Which W0 does the TMA backward kernel effectively use for dx = g0 . W0 ?  (least squares from g0 and gi)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import _lib
from enerf_b200.backends import ffmlp_backend as FB
dev = "cuda"
torch.manual_seed(0)
np.set_printoptions(linewidth=250, precision=3, suppress=True)
B, nl = 128 * 64, 2
nw = 64 * (32 + 64 * (nl - 1) + 16)
w = ((torch.rand(nw, device=dev) * 2 - 1) * (3 / 64) ** 0.5).half()
x = (torch.randn(B, 32, device=dev) * 0.5).half()
g = (torch.randn(B, 16, device=dev) * 0.1).half()
out = torch.empty(B, 16, device=dev, dtype=torch.half)
fb = torch.empty(nl, B, 64, device=dev, dtype=torch.half)
FB.ffmlp_forward(x, w, B, 32, 16, 64, nl, 0, 6, fb, out)
bb = torch.zeros(nl, B, 64, device=dev, dtype=torch.half)
gi_old = torch.zeros(B, 32, device=dev, dtype=torch.half)
gw = torch.zeros(nw, device=dev)
_lib.call("enerf_ffmlp_set_path", 2)
FB.ffmlp_backward(g, x, w, fb, B, 32, 16, 64, nl, 0, 6, True, bb, gi_old, gw)
_lib.call("enerf_ffmlp_set_path", 0)
gi_new = torch.zeros(B, 32, device=dev, dtype=torch.half)
FB.ffmlp_backward(g, x, w, fb, B, 32, 16, 64, nl, 0, 6, True, None, gi_new, gw)
torch.cuda.synchronize()
W0 = w[:64 * 32].float().view(64, 32)
for idx in range(nl):
    g0 = bb[idx].float()
    print(f"bb[{idx}]: |g0 @ W0 - gi_old| = {float((g0 @ W0 - gi_old.float()).abs().max()):.4g}")
g0 = bb[nl - 1].float()
sol = torch.linalg.lstsq(g0.double(), gi_new.double()).solution.float()      # [64, 32]
print("max |W0_eff - W0| =", float((sol - W0).abs().max()))
d = (sol - W0).abs()
print("rows (hidden j) with error > 1e-2:", torch.nonzero(d.max(1).values > 1e-2).flatten().tolist())
print("cols (input n) with error > 1e-2:", torch.nonzero(d.max(0).values > 1e-2).flatten().tolist())
# does W0_eff equal W0 with permuted rows / cols?
best = []
for j in range(64):
    e = (W0 - sol[j:j + 1]).abs().max(1).values
    best.append(int(e.argmin()))
print("W0_eff row j == W0 row:", best)
print("W0_eff[:4,:8]\n", sol[:4, :8].cpu().numpy(), "\nW0[:4,:8]\n", W0[:4, :8].cpu().numpy())
Wh = w[64 * 32:64 * 32 + 4096].float().view(64, 64)
print("W0_eff vs Wh rows? ", float((sol - Wh[:, :32]).abs().max()), float((sol - Wh[:, 32:]).abs().max()))

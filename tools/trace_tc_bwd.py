#!/usr/bin/env python
"""Cycle trace of the recompute-backward kernel (needs tools/bin/libenerf_b200_trace.so, see tools/build_trace.sh):
   ENERF_B200_LIB=tools/bin/libenerf_b200_trace.so python tools/trace_tc_bwd.py [nl]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enerf_b200 import _lib
from enerf_b200.backends import ffmlp_backend as FB

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda"
S = 128 * 148 * 24
nw = 64 * (32 + 64 * (nl - 1) + 16)
w = ((torch.rand(nw, device=dev) * 2 - 1) * (3 / 64) ** 0.5).half()
x = (torch.randn(S, 32, device=dev) * 0.5).half()
g = (torch.randn(S, 16, device=dev) * 0.1).half()
gi = torch.empty(S, 32, device=dev, dtype=torch.half)
gw = torch.empty(nw, device=dev, dtype=torch.float32)
for _ in range(2):
    FB.ffmlp_backward(g, x, w, None, S, 32, 16, 64, nl, 0, 6, True, None, gi, gw)
torch.cuda.synchronize()
buf = torch.zeros(2 * 2 * 4096, dtype=torch.int64, device=dev)
L = _lib.lib()
L.enerf_debug_set_trace.argtypes = [C.c_void_p]
L.enerf_debug_set_trace(buf.data_ptr())
FB.ffmlp_backward(g, x, w, None, S, 32, 16, 64, nl, 0, 6, True, None, gi, gw)
torch.cuda.synchronize()
L.enerf_debug_set_trace(None)
t = buf.cpu().numpy().reshape(-1, 2)
t = t[t[:, 0] > 0]
t = t[np.argsort(t[:, 1])]
t0 = t[0, 1]
T = 2 * (nl - 1) + 3
names = {1: "mma: a_ready seen, issue", 2: "mma: issued+commit", 3: "epi: d_full seen", 4: "epi: math+stores done", 5: "epi: fenced+arrived"}
print(f"{len(t)} events; stages per tile T={T}")
# print tiles 3..5 of slot 0
lines = []
for tag, clk in t:
    kind, st = int(tag) // 1000, int(tag) % 1000
    lines.append((int(clk - t0), kind, st))
prev = None
start = None
count_tiles = 0
for clk, kind, st in lines:
    if kind == 1 and st == 0:
        count_tiles += 1
    if 4 <= count_tiles <= 5:
        print(f"{clk:9d}  (+{clk - (prev if prev is not None else clk):5d})  {names[kind]:28s} stage {st}")
        prev = clk
# summary: mean deltas between consecutive event kinds per stage
import collections
ev = collections.defaultdict(list)
for clk, kind, st in lines:
    ev[(kind, st)].append(clk)
print("mean per-stage intervals (cycles), steady state:")
for st in range(T):
    a = np.array(ev[(1, st)][3:-2]); b = np.array(ev[(2, st)][3:-2]); c = np.array(ev[(3, st)][3:-2]); d = np.array(ev[(4, st)][3:-2]); e = np.array(ev[(5, st)][3:-2])
    n = min(len(a), len(b), len(c), len(d), len(e))
    if n == 0:
        continue
    a, b, c, d, e = a[:n], b[:n], c[:n], d[:n], e[:n]
    nxt = np.array(ev[(1, (st + 1) % T)][3:-2])
    print(f"  stage {st}: issue {np.mean(b - a):6.0f} | commit->epi sees d_full {np.mean(c - b):6.0f} | epi math {np.mean(d - c):6.0f} | fence+arrive {np.mean(e - d):6.0f}")
tile_starts = np.array(ev[(1, 0)])
print("cycles per tile (slot 0):", np.mean(np.diff(tile_starts)[2:-2]))

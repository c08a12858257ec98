"""Times the fused field forwards / backwards alone on N marcher-sized rows (CUDA events, L2 flushed between launches).
`ENERF_B200_LIB=<variant .so> python tools/mlp_probe.py` compares builds (e.g. -DENERF_FWD_SLOTS=5)."""
import json
import sys

import torch

from enerf_b200 import _lib
from enerf_b200._lib import ptr, stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3286912          # the bench step's sample count (a multiple of 128)
dev = torch.device("cuda:0")
torch.manual_seed(0)
feat = (torch.randn(N, 32, device=dev) * 0.3).half()
dirs = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1)
ws = (torch.randn(64 * (32 + 64 + 16), device=dev) * 0.15).half()
wc = (torch.randn(64 * (32 + 128 + 16), device=dev) * 0.15).half()
sigma = torch.empty(N, device=dev)
cin = torch.empty(N, 32, dtype=torch.float16, device=dev)
rgb = torch.empty(N, 1, device=dev)
g_sigma, g_rgb = torch.randn(N, device=dev), torch.randn(N, 1, device=dev)
dcin, dfeat = torch.empty_like(cin), torch.empty_like(feat)
dws, dwc = torch.zeros(ws.numel(), device=dev), torch.zeros(wc.numel(), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[reps // 2]


out = {"rows": N, "lib": _lib.LIB_PATH}
out["sigma_forward_ms"] = timed(lambda: _lib.call("enerf_field_sigma_forward", ptr(feat), ptr(ws), ptr(dirs), N, 2, None, ptr(sigma), ptr(cin), stream()))
out["color_forward_ms"] = timed(lambda: _lib.call("enerf_field_color_forward", ptr(cin), ptr(wc), N, 3, 1, None, ptr(rgb), None, stream()))
out["density_forward_ms"] = timed(lambda: _lib.call("enerf_field_density_forward", ptr(feat), ptr(ws), N, 2, ptr(sigma), None, stream()))
out["color_backward_ms"] = timed(lambda: _lib.call("enerf_field_color_backward", ptr(g_rgb), ptr(rgb), 1, ptr(cin), ptr(wc), None, N, 3, ptr(dcin), ptr(dwc),
                                                   None, stream()))
out["sigma_backward_ms"] = timed(lambda: _lib.call("enerf_field_sigma_backward", ptr(g_sigma), ptr(sigma), ptr(dcin), ptr(feat), ptr(ws), None, N, 2,
                                                   ptr(dfeat), ptr(dws), stream()))
out["checksum_bwd"] = [float(dcin.double().abs().sum()), float(dfeat.double().abs().sum()), float(dwc.double().abs().sum()), float(dws.double().abs().sum())]
out["checksum"] = [float(sigma.double().sum()), float(rgb.double().sum()), float(cin.double().abs().sum())]
print(json.dumps(out))

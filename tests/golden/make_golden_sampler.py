#!/usr/bin/env python
"""Golden vectors for the event-pair sampler, produced by the REFERENCE's own `EventNeRFDataset.collate`
(nerf/provider.py:1364-1448, accumulate_evs branch): its source is extracted with `ast` and executed on the CPU with a stand-in
`self` that carries one synthetic event frame, while `np.random.randint` is scripted so that the draws are the integers the
device sampler derives from the stored uniform variates (start = floor(u*E), end = low + floor(u*(high-low)), fp32).
Build container only (needs /root/reference).  Writes tests/golden/sampler.npz.
"""
import ast
import os
import types

import numpy as np
import torch

from make_golden_events import REF, HERE, load_functions


def load_method(path, cls, name, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == name:
                    exec(compile(ast.Module(body=[item], type_ignores=[]), path, "exec"), ns)
                    return ns[name]
    raise KeyError(name)


def synthetic_frame(rng, n_pixels=300, acc_max=5):
    """events grouped by pixel (>= 2 per pixel, time-sorted inside a pixel) + the successor tables of provider.py:1166-1187"""
    counts = rng.integers(2, 13, n_pixels)
    pix = rng.choice(346 * 260, n_pixels, replace=False)
    ev = []
    for c, p in zip(counts, pix):
        ts = np.sort(rng.random(c)) * 1e6
        for t in ts:
            ev.append((p % 346, p // 346, t, rng.choice([-1.0, 1.0])))
    events = np.asarray(ev, np.float32)
    cum = np.cumsum(counts)                                  # provider.py:1174
    num_evs = int(cum[-1])                                   # :1175
    idx_no_successor = cum - 1                               # :1178
    num_succ = np.zeros(num_evs, np.int64)                   # :1181-1186
    j = 0
    for i in range(num_evs):
        if i >= cum[j]:
            j += 1
        num_succ[i] = cum[j] - i - 1
    return events, num_evs, idx_no_successor, num_succ


def main():
    u = load_functions(os.path.join(REF, "nerf", "utils.py"), ["custom_meshgrid", "get_rays", "get_event_rays"])
    ns = {"np": np, "torch": torch, "get_event_rays": u["get_event_rays"], "get_rays": u["get_rays"]}
    collate = load_method(os.path.join(REF, "nerf", "provider.py"), "EventNeRFDataset", "collate", ns)
    rng = np.random.default_rng(99)
    out = {}
    for case, acc_max in (("a", 5), ("b", 0)):
        events, E, no_succ, num_succ = synthetic_frame(rng, acc_max=acc_max)
        M = 512
        g = torch.Generator().manual_seed(5 + acc_max)
        q = torch.linalg.qr(torch.randn(E, 3, 3, generator=g))[0]
        poses_evs = torch.cat([q, torch.randn(E, 3, 1, generator=g) * 0.2], dim=-1)          # [E,3,4]
        u_start = rng.random(M).astype(np.float32)
        u_end = rng.random(M).astype(np.float32)
        u_start[:4] = [0.0, 0.99999994, (no_succ[3] + 0.5) / E, (no_succ[0] + 0.5) / E]        # edges: first / last event, events without successor
        calls = {"n": 0}

        def scripted_randint(low, high=None, size=None):
            k = calls["n"]
            calls["n"] += 1
            if k == 0:                                                                          # provider.py:1369
                assert low == 0 and high == E and size == M
                return np.minimum((u_start * np.float32(E)).astype(np.int64), E - 1)
            span = high - low                                                                   # provider.py:1382
            return np.array([low + min(int(np.float32(u_end[k - 1]) * np.float32(span)), span - 1)])

        fake = types.SimpleNamespace(
            frame_idxs=[0], accumulate_evs=1, num_evs={0: E}, batch_size_evs=M, idx_no_successor={0: no_succ}, num_successor_evs={0: num_succ},
            acc_max_num_evs=acc_max, events={0: torch.from_numpy(events)}, precompute_evs_poses=True, poses_evs={0: poses_evs},
            intrinsics_evs=np.array([250.1, 249.7, 172.4, 131.9], np.float32), poses=torch.eye(4)[None], device="cpu", error_map=None,
            intrinsics=np.array([250.1, 249.7, 172.4, 131.9], np.float32), H=260, W=346, num_rays=16, negative_event_sampling=0, images=None,
            training=True)
        keep = np.random.randint
        np.random.randint = scripted_randint
        try:
            res = collate(fake, [0])
        finally:
            np.random.randint = keep
        assert calls["n"] == M + 1
        out.update({f"{case}_events": events, f"{case}_no_succ": no_succ, f"{case}_num_succ": num_succ, f"{case}_acc_max": np.int64(acc_max),
                    f"{case}_poses_evs": poses_evs.numpy(), f"{case}_u_start": u_start, f"{case}_u_end": u_end, f"{case}_intr": fake.intrinsics_evs,
                    f"{case}_pols": res["pols"].numpy(), f"{case}_o1": res["rays_evs_o1"].numpy(), f"{case}_d1": res["rays_evs_d1"].numpy(),
                    f"{case}_o2": res["rays_evs_o2"].numpy(), f"{case}_d2": res["rays_evs_d2"].numpy()})
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("wrote sampler.npz:", os.path.getsize(os.path.join(HERE, "sampler.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

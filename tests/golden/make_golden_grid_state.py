#!/usr/bin/env python
"""Golden vectors for the occupancy-grid maintenance, produced by the REFERENCE's own Python:
`NeRFRenderer.mark_untrained_grid` and `NeRFRenderer.update_extra_state` (nerf/renderer.py:408-563), imported unchanged through
oracle/ref_python.py and executed on the CPU with
  * a scripted `density()` built from IEEE-exact torch ops (oracle/grid_state.py), so CPU and GPU values are bit-identical,
  * `torch.rand_like` / `torch.randint` patched to hash-based draws the GPU test can regenerate,
  * `raymarching.morton3D(_invert)` / `packbits` served by the C oracle.
Sequence: mark_untrained_grid -> 2 full refreshes -> (iter_density := 16) -> 2 partial refreshes.
Build container only (needs /root/reference).  Writes tests/golden/grid_state.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import grid_state as gs  # noqa: E402
from oracle import oracle, ref_python  # noqa: E402

BOUND, H = 2, 128


def main():
    # one thread: `tmp_grid[cas, indices] = sigmas` (renderer.py:545) has duplicate indices in the partial branch, and torch's CPU index_put
    # only resolves them deterministically ("the last one wins") when it runs sequentially
    torch.set_num_threads(1)
    ns = ref_python.load()
    r = ns.NeRFRenderer(bound=BOUND, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1)
    C, cells = r.cascade, H ** 3
    r.density = lambda x: {'sigma': gs.scripted_density(x)}
    poses = gs.camera_ring(7, 1.1 * BOUND)
    intrinsic = np.array([410.0, 395.0, 160.0, 120.0], np.float32)          # fx, fy, cx, cy: narrow enough to leave cells unseen
    out = {"bound": np.int64(BOUND), "poses": poses, "intrinsic": intrinsic, "decay": np.float32(0.95)}

    r.mark_untrained_grid(torch.from_numpy(poses), intrinsic)
    g = r.density_grid.numpy()
    out["untrained_bits"] = np.packbits((g.reshape(-1) < 0).astype(np.uint8), bitorder="little")
    print("untrained cells:", int((g < 0).sum()), "of", g.size)

    state = {"upd": 0, "cas": 0, "phase": 0}
    real_rand_like, real_randint = torch.rand_like, torch.randint

    def full_rand_like(t):
        # one call per cascade, rows in meshgrid order (x slowest); the same cell gets the same variates in any order
        n = t.shape[0]
        lin = np.arange(n, dtype=np.int64)
        coords = np.stack([lin // (H * H), (lin // H) % H, lin % H], axis=-1).astype(np.int32)
        m = oracle.morton3D(coords)
        u = gs.full_noise(state["upd"], state["cas"], m)
        state["cas"] += 1
        return torch.from_numpy(u)

    def partial_randint(low, high, size, **kw):
        cas = state["cas"]
        if state["phase"] == 0:                                # coords = torch.randint(0, H, (N, 3))
            assert (low, high) == (0, H) and tuple(size) == (cells // 4, 3)
            n_occ = int((r.density_grid[cas] > 0).sum())
            state["draws"] = gs.partial_draws(state["upd"], cas, cells // 4, H, n_occ)
            state["phase"] = 1
            return torch.from_numpy(state["draws"][0])
        assert low == 0 and list(size) == [cells // 4]         # rand_mask = torch.randint(0, Nz, [N])
        assert high == int((r.density_grid[cas] > 0).sum())
        state["phase"] = 2
        return torch.from_numpy(state["draws"][1])

    def partial_rand_like(t):
        assert state["phase"] == 2 and t.shape == (cells // 2, 3)
        u = state["draws"][2]
        state["phase"] = 0
        state["cas"] += 1
        return torch.from_numpy(u)

    r.step_counter[:3, 0] = torch.tensor([1000, 1300, 1100], dtype=torch.int32)
    r.local_step = 3
    try:
        for upd in range(2):
            state.update(upd=upd, cas=0)
            torch.rand_like = full_rand_like
            r.update_extra_state()
            torch.rand_like = real_rand_like
            for k, v in gs.snapshot(r.density_grid.numpy(), r.density_bitfield.numpy()).items():
                out[f"full{upd}_{k}"] = v
            out[f"full{upd}_mean_density"] = np.float64(r.mean_density)
            print(f"full {upd}: mean_density {r.mean_density:.6f}, occupied bits {int(np.unpackbits(r.density_bitfield.numpy()).sum())}")
        out["mean_count"] = np.int64(r.mean_count)
        r.iter_density = 16
        for upd in range(2):
            state.update(upd=upd, cas=0, phase=0)
            torch.rand_like, torch.randint = partial_rand_like, partial_randint
            r.update_extra_state()
            torch.rand_like, torch.randint = real_rand_like, real_randint
            for k, v in gs.snapshot(r.density_grid.numpy(), r.density_bitfield.numpy()).items():
                out[f"partial{upd}_{k}"] = v
            out[f"partial{upd}_mean_density"] = np.float64(r.mean_density)
            print(f"partial {upd}: mean_density {r.mean_density:.6f}, occupied bits {int(np.unpackbits(r.density_bitfield.numpy()).sum())}")
    finally:
        torch.rand_like, torch.randint = real_rand_like, real_randint
    path = os.path.join(HERE, "grid_state.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

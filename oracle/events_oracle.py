"""CPU restatement (numpy, fp32 arithmetic in the reference's order of operations) of the functions next to the hot path
(SURVEY.md §8f N1, N2).  TEST INFRASTRUCTURE ONLY — imported by tests/ alone; pinned against tests/golden/events.npz, which
the reference's own Python code produced (tests/golden/make_golden_events.py).
"""
import numpy as np

F = np.float32


def pixel_dirs(px, py, intrinsics):
    """nerf/utils.py:160-165 and :203-207: normalize(((x-cx)/fx, (y-cy)/fy, 1))"""
    fx, fy, cx, cy = (F(v) for v in intrinsics)
    x = (px.astype(F) - cx) / fx
    y = (py.astype(F) - cy) / fy
    z = np.ones_like(x)
    d = np.stack([x, y, z], -1)
    return d / np.sqrt((d * d).sum(-1, keepdims=True, dtype=F))


def get_rays(poses, intrinsics, H, W, inds=None):
    """nerf/utils.py:110-169.  poses [B,4,4]; inds [N] flat pixel indices (None: all).  -> rays_o, rays_d [B,N,3]"""
    if inds is None:
        inds = np.arange(H * W)
    i, j = (inds % W).astype(F), (inds // W).astype(F)
    d = pixel_dirs(i, j, intrinsics)                                           # [N,3]
    rays_d = np.einsum("nk,bik->bni", d, poses[:, :3, :3].astype(F)).astype(F)  # directions @ R^T  (utils.py:166)
    rays_o = np.broadcast_to(poses[:, None, :3, 3], rays_d.shape).astype(F)
    return rays_o, rays_d


def get_event_rays(xs, ys, c2w_before, c2w_at, intrinsics):
    """nerf/utils.py:185-216.  poses [...,N,3,4] -> dict of [...,N,3]"""
    d = pixel_dirs(xs, ys, intrinsics)
    out = {}
    for tag, P in (("1", c2w_before), ("2", c2w_at)):
        out["rays_evs_o" + tag] = P[..., :3, 3].astype(F)
        out["rays_evs_d" + tag] = (d[..., None, :] * P[..., :3, :3].astype(F)).sum(-1, dtype=F)
    return out


def rgb_to_luma(rgb):
    """utils/event_utils.py:23-53, esim coefficients"""
    return (rgb.astype(F) * np.array([0.299, 0.587, 0.114], F)).sum(-1, keepdims=True, dtype=F)


def lin_log(color, thres=20):
    """utils/event_utils.py:55-66"""
    slope = F(np.log(thres) / thres)
    c = color.astype(F)
    return np.where(c < thres, slope * c, np.log(np.maximum(c, F(1e-30)))).astype(F)


def event_loss(img1, img2, pols, use_luma, linlog, C_thres, event_only, log_thres=20.0):
    """nerf/utils.py:494-528 -> (loss float64, delta_linlog [..,N,C']).  Evaluated in float64 after the fp32 log-intensities."""
    l1, l2 = (rgb_to_luma(img1), rgb_to_luma(img2)) if use_luma else (img1.astype(F), img2.astype(F))
    if linlog:
        p1, p2 = lin_log(l1 * F(255)), lin_log(l2 * F(255))
    else:
        p1, p2 = np.log(np.maximum(l1 * F(255), F(log_thres))), np.log(np.maximum(l2 * F(255), F(log_thres)))
    delta = (p2 - p1).astype(F)
    gt = pols[..., None].astype(np.float64)
    d64 = delta.astype(np.float64)
    if C_thres != -1:
        return float(np.mean((d64 - gt * C_thres) ** 2)), delta
    w = 20.0 * (1.0 if event_only else 20.0)
    dn = d64 / (np.linalg.norm(d64, axis=-2, keepdims=True) + 1e-9)
    pn = gt / (np.linalg.norm(gt, axis=-2, keepdims=True) + 1e-9)
    return float(w * np.mean((dn - pn) ** 2)), delta


def sample_event_pairs(events, num_succ, idx_no_successor, acc_max, u_start, u_end):
    """nerf/provider.py:1364-1405 (accumulate_evs branch) with the integer draws derived from uniform variates the way the
    device sampler does (fp32: start = floor(u*E), end = start+1+floor(u*n)).  -> eidx, eidx_end, pols, xs, ys"""
    E = events.shape[0]
    no_succ = set(int(i) for i in idx_no_successor)
    eidx = np.minimum((u_start.astype(F) * F(E)).astype(np.int64), E - 1)
    eidx = np.asarray([i - 1 if int(i) in no_succ else i for i in eidx])
    ends, pols = [], []
    for k, s in enumerate(eidx):
        n = int(num_succ[s])
        if acc_max:
            n = min(n, acc_max + 1)
        e = s + 1 + min(int(F(u_end[k]) * F(n)), n - 1)
        pols.append(events[s + 1:e + 1, 3].sum())
        ends.append(e)
    return eidx, np.asarray(ends), np.asarray(pols, F), events[eidx, 0], events[eidx, 1]

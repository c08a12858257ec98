"""The steps on either side of the volume-rendering path, fused — host-side mirrors of the
reference functions they replace, same names, argument meaning and return values:

  get_rays(poses, intrinsics, H, W, N=-1, error_map=None)      nerf/utils.py:110-169
  get_event_rays(xs, ys, c2w_before, c2w_at, intrinsics)       nerf/utils.py:185-216
  event_loss(image1, image2, pols, ...)                        nerf/utils.py:494-528 (the body of train_step_events between
                                                               the two renders and `loss = loss_evs`)

`*_with_near_far` variants additionally return the near/far of raymarching.near_far_from_aabb computed in the same kernel.
No CPU fallback: CUDA tensors only.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import need_cuda, ptr, stream


def _intr(intrinsics):
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    return fx, fy, cx, cy


def get_rays_with_near_far(poses, intrinsics, H, W, N=-1, error_map=None, aabb=None, min_near=0.2):
    need_cuda(poses)
    dev = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = _intr(intrinsics)
    results = {}
    inds, per_pose = None, 0
    if N > 0:
        N = min(N, H * W)
        if error_map is None:
            inds = torch.randint(0, H * W, size=[N], device=dev)          # may duplicate (utils.py:137); shared by all poses
            results['inds'] = inds.expand([B, N])
        else:
            # importance sampling on the 128x128 error map, then a random pixel inside the chosen cell (utils.py:140-150)
            inds_coarse = torch.multinomial(error_map.to(dev), N, replacement=False)         # [B, N] in [0, 128*128)
            cell_x, cell_y = inds_coarse // 128, inds_coarse % 128
            sx, sy = H / 128, W / 128
            px = (cell_x * sx + torch.rand(B, N, device=dev) * sx).long().clamp(max=H - 1)
            py = (cell_y * sy + torch.rand(B, N, device=dev) * sy).long().clamp(max=W - 1)
            inds, per_pose = (px * W + py).contiguous(), 1
            results['inds_coarse'] = inds_coarse
            results['inds'] = inds
    n = N if N > 0 else H * W
    P = poses.detach().float().contiguous()
    rays_o = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    nears = fars = None
    a = None
    if aabb is not None:
        a = aabb.detach().float().contiguous()
        nears = torch.empty(B, n, dtype=torch.float32, device=dev)
        fars = torch.empty(B, n, dtype=torch.float32, device=dev)
    _lib.call("enerf_get_rays", ptr(P), fx, fy, cx, cy, H, W, ptr(inds), per_pose, B, n, ptr(a), float(min_near), ptr(rays_o), ptr(rays_d), ptr(nears),
              ptr(fars), stream())
    results['rays_o'] = rays_o
    results['rays_d'] = rays_d
    if aabb is not None:
        results['nears'], results['fars'] = nears, fars
    return results


def get_rays(poses, intrinsics, H, W, N=-1, error_map=None):
    return get_rays_with_near_far(poses, intrinsics, H, W, N, error_map)


def get_event_rays_with_near_far(xs, ys, c2w_before, c2w_at, intrinsics, aabb=None, min_near=0.2):
    need_cuda(xs, ys, c2w_before, c2w_at)
    dev = xs.device
    fx, fy, cx, cy = _intr(intrinsics)
    lead = c2w_before.shape[:-2]                                 # (B, Nevs) in the reference, B == 1
    n = xs.numel()
    if c2w_before.numel() != n * 12:
        raise RuntimeError("get_event_rays: one [3,4] pose per event expected")
    x = xs.detach().float().contiguous().view(-1)
    y = ys.detach().float().contiguous().view(-1)
    pb = c2w_before.detach().float().contiguous().view(-1, 3, 4)
    pa = c2w_at.detach().float().contiguous().view(-1, 3, 4)
    o1, d1, o2, d2 = (torch.empty(n, 3, dtype=torch.float32, device=dev) for _ in range(4))
    nf1 = nf2 = a = None
    if aabb is not None:
        a = aabb.detach().float().contiguous()
        nf1 = torch.empty(2, n, dtype=torch.float32, device=dev)
        nf2 = torch.empty(2, n, dtype=torch.float32, device=dev)
    _lib.call("enerf_event_rays", ptr(x), ptr(y), ptr(pb), ptr(pa), fx, fy, cx, cy, n, ptr(a), float(min_near), ptr(o1), ptr(d1), ptr(o2), ptr(d2),
              ptr(nf1), ptr(nf2), stream())
    out = {"rays_evs_o1": o1.view(*lead, 3), "rays_evs_d1": d1.view(*lead, 3), "rays_evs_o2": o2.view(*lead, 3), "rays_evs_d2": d2.view(*lead, 3)}
    if aabb is not None:
        out.update(nears1=nf1[0].view(*lead), fars1=nf1[1].view(*lead), nears2=nf2[0].view(*lead), fars2=nf2[1].view(*lead))
    return out


def get_event_rays(xs, ys, c2w_before, c2w_at, intrinsics):
    return get_event_rays_with_near_far(xs, ys, c2w_before, c2w_at, intrinsics)


class _EventLoss(Function):
    @staticmethod
    def forward(ctx, img1, img2, pols, use_luma, linlog, log_thres, c_thres, weight):
        need_cuda(img1, img2, pols)
        C = img1.shape[-1]
        a = img1.detach().float().contiguous().view(-1, C)
        b = img2.detach().float().contiguous().view(-1, C)
        p = pols.detach().float().contiguous().view(-1)
        N = a.shape[0]
        if b.shape != a.shape or p.shape[0] != N:
            raise RuntimeError("event_loss: image1, image2 [.., N, C] and pols [.., N] must agree")
        Cp = 1 if use_luma else C
        dev = a.device
        delta = torch.empty(N, Cp, dtype=torch.float32, device=dev)
        acc = torch.empty(16, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.call("enerf_event_loss_forward", ptr(a), ptr(b), ptr(p), N, C, int(use_luma), int(linlog), float(log_thres), float(c_thres), float(weight),
                  ptr(delta), ptr(acc), ptr(loss), stream())
        ctx.save_for_backward(a, b, p, delta, acc)
        ctx.meta = (N, C, int(use_luma), int(linlog), float(log_thres), float(c_thres), float(weight), img1.shape, img2.shape)
        ctx.mark_non_differentiable(delta)
        return loss.view(()), delta

    @staticmethod
    def backward(ctx, g_loss, _g_delta):
        a, b, p, delta, acc = ctx.saved_tensors
        N, C, use_luma, linlog, log_thres, c_thres, weight, s1, s2 = ctx.meta
        g = g_loss.detach().float().contiguous().view(1)
        g1 = torch.empty_like(a)
        g2 = torch.empty_like(b)
        _lib.call("enerf_event_loss_backward", ptr(a), ptr(b), ptr(p), ptr(delta), ptr(acc), ptr(g), N, C, use_luma, linlog, log_thres, c_thres, weight,
                  ptr(g1), ptr(g2), stream())
        return g1.view(s1), g2.view(s2), None, None, None, None, None, None


def event_loss(image1, image2, pols, use_luma=False, linlog=True, C_thres=-1, event_only=True, log_thres=20.0):
    """loss_evs and delta_linlog of nerf/utils.py:494-528 for the renders `image1`, `image2` [B,N,C] of the two poses of each
    event pair and the accumulated polarities `pols` [B,N] (B == 1).  Returns (loss, delta_linlog [B,N,C'])."""
    if C_thres != -1:
        weight = 1.0
    else:
        weight = 20.0 * (1.0 if event_only else 20.0)      # utils.py:522-525
    loss, delta = _EventLoss.apply(image1, image2, pols, bool(use_luma), bool(linlog), float(log_thres), float(C_thres), weight)
    return loss, delta.view(*image1.shape[:-1], delta.shape[-1])


class EventPairSampler:
    """Device-resident replacement for the per-pair Python loop of `EventNeRFDataset.collate` (nerf/provider.py:1364-1405,
    accumulate_evs branch) for ONE event frame: keeps the frame's events, their polarity prefix sum, the successor counts and the
    "last event of its pixel" mask on the GPU and draws M pairs per call with one kernel.

        sampler = EventPairSampler(events_fidx, num_successor_evs_fidx, idx_no_successor_fidx, acc_max_num_evs, poses_evs_fidx)
        batch = sampler.sample(M, intrinsics_evs)      # -> xs, ys, pols [1,M], eidx, eidx_end, rays_evs_o1/d1/o2/d2 [1,M,3]

    The pairs follow the reference's distribution (uniform start event, predecessor if it has no successor, uniform end event among
    the at most acc_max_num_evs+1 successors); the random variates come from torch's CUDA generator instead of NumPy's global one.
    """

    def __init__(self, events, num_successor_evs, idx_no_successor, acc_max_num_evs=0, poses_evs=None, device="cuda"):
        self.events = torch.as_tensor(events, dtype=torch.float32).to(device).contiguous()
        E = self.events.shape[0]
        self.E = E
        pol = self.events[:, 3].double()
        self.pol_prefix = torch.cat([torch.zeros(1, dtype=torch.float64, device=device), torch.cumsum(pol, 0)]).contiguous()
        self.num_succ = torch.as_tensor(num_successor_evs).to(device=device, dtype=torch.int32).contiguous()
        mask = torch.zeros(E, dtype=torch.uint8, device=device)
        mask[torch.as_tensor(idx_no_successor).to(device=device, dtype=torch.long)] = 1
        self.no_succ = mask
        # The kernel steps back one event where a pixel's last event was drawn and then draws among `num_successor` followers.
        # The reference's provider guarantees the layout that makes this safe (pixels with a single event are dropped,
        # provider.py:1164); arbitrary arrays are checked here, once, instead of reading out of bounds on the device.
        if E < 2:
            raise ValueError("EventPairSampler: a frame needs at least two events")
        if self.num_succ.shape[0] != E:
            raise ValueError("EventPairSampler: num_successor_evs must have one entry per event")
        if bool(mask[0]):
            raise ValueError("EventPairSampler: event 0 is marked as having no successor (it has no predecessor to fall back to)")
        has_succ = mask == 0
        idx = torch.arange(E, device=device)
        if bool(((self.num_succ < 1) & has_succ).any()):
            raise ValueError("EventPairSampler: an event that is not the last of its pixel must have num_successor_evs >= 1")
        if bool(((idx + self.num_succ.long() >= E) & has_succ).any()):
            raise ValueError("EventPairSampler: num_successor_evs points past the end of the frame")
        if bool((mask[1:].bool() & mask[:-1].bool()).any()):
            raise ValueError("EventPairSampler: two consecutive events without successor (a pixel with a single event)")
        self.acc_max = int(acc_max_num_evs or 0)
        self.poses_evs = None if poses_evs is None else torch.as_tensor(poses_evs, dtype=torch.float32).to(device).contiguous()

    def sample(self, M, intrinsics_evs=None, u_start=None, u_end=None, aabb=None, min_near=0.2):
        dev = self.events.device
        u_start = torch.rand(M, device=dev) if u_start is None else u_start.to(dev).float().contiguous()
        u_end = torch.rand(M, device=dev) if u_end is None else u_end.to(dev).float().contiguous()
        eidx = torch.empty(M, dtype=torch.int64, device=dev)
        eidx_end = torch.empty(M, dtype=torch.int64, device=dev)
        pols, xs, ys = (torch.empty(M, dtype=torch.float32, device=dev) for _ in range(3))
        _lib.call("enerf_sample_event_pairs", ptr(self.events), ptr(self.pol_prefix), ptr(self.num_succ), ptr(self.no_succ), self.E, M, self.acc_max,
                  ptr(u_start), ptr(u_end), ptr(eidx), ptr(eidx_end), ptr(pols), ptr(xs), ptr(ys), stream())
        out = {"eidx": eidx, "eidx_end": eidx_end, "pols": pols.unsqueeze(0), "xs": xs.unsqueeze(0), "ys": ys.unsqueeze(0)}
        if self.poses_evs is not None and intrinsics_evs is not None:
            p1 = self.poses_evs[eidx].unsqueeze(0)          # (1, M, 3, 4), provider.py:1417-1418
            p2 = self.poses_evs[eidx_end].unsqueeze(0)
            out.update(get_event_rays_with_near_far(out["xs"], out["ys"], p1, p2, intrinsics_evs, aabb=aabb, min_near=min_near))
        return out

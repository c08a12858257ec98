"""The next rows of the path: device ray generation and the fused event-loss tail.
CPU: the numpy oracle reproduces the golden vectors the reference's own Python code produced (tests/golden/events.npz).
GPU: the kernels reproduce the same vectors (forward, gradients, fused near/far)."""
import os

import numpy as np
import pytest
import torch

from oracle import events_oracle as eo
from oracle import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "events.npz"))
HAS_GPU = torch.cuda.is_available()


def _cases():
    return [(int(a), int(b), float(c), int(d)) for a, b, c, d in G["el_cases"]]


def _tag(use_luma, linlog, C_thres, event_only):
    return f"el_{use_luma}{linlog}{int(C_thres != -1)}{event_only}"


# ------------------------------------------------------------------ CPU: oracle pinned on the reference's outputs
def test_oracle_get_rays_reproduces_reference():
    H, W = (int(v) for v in G["gr_hw"])
    o, d = eo.get_rays(G["gr_poses"], G["gr_intr"], H, W)
    assert np.array_equal(o, G["gr_o_all"]) and np.allclose(d, G["gr_d_all"], atol=2e-7)
    o, d = eo.get_rays(G["gr_poses"], G["gr_intr"], H, W, G["gr_inds"])
    assert np.array_equal(o, G["gr_o_sel"]) and np.allclose(d, G["gr_d_sel"], atol=2e-7)


def test_oracle_get_event_rays_reproduces_reference():
    r = eo.get_event_rays(G["er_xs"], G["er_ys"], G["er_pb"], G["er_pa"], G["er_intr"])
    for k in ("o1", "d1", "o2", "d2"):
        assert np.allclose(r["rays_evs_" + k], G["er_" + k], atol=2e-7), k


def test_oracle_event_loss_reproduces_reference():
    for use_luma, linlog, C_thres, event_only in _cases():
        tag = _tag(use_luma, linlog, C_thres, event_only)
        loss, delta = eo.event_loss(G["el_img1"], G["el_img2"], G["el_pols"], use_luma, linlog, C_thres, event_only)
        assert np.allclose(delta, G[tag + "_delta"], atol=2e-6), tag
        assert abs(loss - float(G[tag + "_loss"])) <= 2e-5 * abs(float(G[tag + "_loss"])) + 1e-9, (tag, loss, float(G[tag + "_loss"]))


# ------------------------------------------------------------------ GPU: kernels vs the same vectors
@pytest.mark.gpu
def test_gpu_get_rays_and_fused_near_far():
    from enerf_b200 import events
    dev = "cuda"
    H, W = (int(v) for v in G["gr_hw"])
    poses = torch.from_numpy(G["gr_poses"]).to(dev)
    r = events.get_rays(poses, G["gr_intr"], H, W, -1)
    assert torch.equal(r["rays_o"].cpu(), torch.from_numpy(G["gr_o_all"]))
    assert np.allclose(r["rays_d"].cpu().numpy(), G["gr_d_all"], atol=3e-7)
    # chosen pixels + the slab test in the same kernel vs the stand-alone near_far kernel / oracle
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1], device=dev)
    from enerf_b200 import _lib
    from enerf_b200._lib import ptr, stream
    inds = torch.from_numpy(G["gr_inds"]).to(dev)
    B, N = 3, inds.numel()
    ro, rd = torch.empty(B, N, 3, device=dev), torch.empty(B, N, 3, device=dev)
    nears, fars = torch.empty(B, N, device=dev), torch.empty(B, N, device=dev)
    fx, fy, cx, cy = (float(v) for v in G["gr_intr"])
    _lib.call("enerf_get_rays", ptr(poses.contiguous()), fx, fy, cx, cy, H, W, ptr(inds), 0, B, N, ptr(aabb), 0.2, ptr(ro), ptr(rd), ptr(nears), ptr(fars), stream())
    assert np.allclose(rd.cpu().numpy(), G["gr_d_sel"], atol=3e-7) and torch.equal(ro.cpu(), torch.from_numpy(G["gr_o_sel"]))
    wn, wf = oracle.near_far_from_aabb(ro.cpu().numpy().reshape(-1, 3), rd.cpu().numpy().reshape(-1, 3), aabb.cpu().numpy(), 0.2)
    assert np.array_equal(nears.cpu().numpy().reshape(-1), wn) and np.array_equal(fars.cpu().numpy().reshape(-1), wf)
    # sampled variant returns indices in range and consistent rays
    torch.manual_seed(0)
    r = events.get_rays(poses, G["gr_intr"], H, W, 100)
    o2, d2 = eo.get_rays(G["gr_poses"], G["gr_intr"], H, W, r["inds"][0].cpu().numpy())
    assert r["inds"].shape == (3, 100) and np.allclose(r["rays_d"].cpu().numpy(), d2, atol=3e-7)
    # error-map importance sampling: per-pose pixel sets
    emap = torch.rand(3, 128 * 128, device=dev)
    r = events.get_rays(poses, G["gr_intr"], H, W, 64, error_map=emap)
    assert r["inds"].shape == (3, 64) and r["inds_coarse"].shape == (3, 64) and int(r["inds"].max()) < H * W
    for b in range(3):
        _, db = eo.get_rays(G["gr_poses"][b:b + 1], G["gr_intr"], H, W, r["inds"][b].cpu().numpy())
        assert np.allclose(r["rays_d"][b].cpu().numpy(), db[0], atol=3e-7)


@pytest.mark.gpu
def test_gpu_get_event_rays():
    from enerf_b200 import events
    dev = "cuda"
    t = lambda k: torch.from_numpy(G[k]).to(dev)          # noqa: E731
    aabb = torch.tensor([-2.0, -2, -2, 2, 2, 2], device=dev)
    r = events.get_event_rays_with_near_far(t("er_xs"), t("er_ys"), t("er_pb"), t("er_pa"), G["er_intr"], aabb=aabb, min_near=0.2)
    for k in ("o1", "d1", "o2", "d2"):
        got = r["rays_evs_" + k].cpu().numpy()
        assert got.shape == G["er_" + k].shape and np.allclose(got, G["er_" + k], atol=3e-7), k
    for v in ("1", "2"):
        wn, wf = oracle.near_far_from_aabb(r["rays_evs_o" + v].cpu().numpy().reshape(-1, 3), r["rays_evs_d" + v].cpu().numpy().reshape(-1, 3),
                                           aabb.cpu().numpy(), 0.2)
        assert np.array_equal(r["nears" + v].cpu().numpy().reshape(-1), wn) and np.array_equal(r["fars" + v].cpu().numpy().reshape(-1), wf)
    plain = events.get_event_rays(t("er_xs"), t("er_ys"), t("er_pb"), t("er_pa"), G["er_intr"])
    assert set(plain) == {"rays_evs_o1", "rays_evs_d1", "rays_evs_o2", "rays_evs_d2"}


@pytest.mark.gpu
def test_gpu_event_loss_forward_backward_matches_reference():
    from enerf_b200 import events
    dev = "cuda"
    for use_luma, linlog, C_thres, event_only in _cases():
        tag = _tag(use_luma, linlog, C_thres, event_only)
        a = torch.from_numpy(G["el_img1"]).to(dev).requires_grad_(True)
        b = torch.from_numpy(G["el_img2"]).to(dev).requires_grad_(True)
        p = torch.from_numpy(G["el_pols"]).to(dev)
        loss, delta = events.event_loss(a, b, p, use_luma=use_luma, linlog=linlog, C_thres=C_thres, event_only=event_only)
        want = float(G[tag + "_loss"])
        assert abs(float(loss) - want) <= 2e-4 * abs(want) + 1e-8, (tag, float(loss), want)
        assert np.allclose(delta.cpu().numpy(), G[tag + "_delta"], atol=5e-6), tag
        (loss * 3.0).backward()
        for got, key in ((a.grad, "_g1"), (b.grad, "_g2")):
            w = 3.0 * G[tag + key]
            err = np.abs(got.cpu().numpy() - w).max()
            assert err <= 2e-4 * np.abs(w).max() + 1e-9, (tag, key, err, np.abs(w).max())
    # one-channel images (what every shipped config trains: out_dim_color = 1)
    a1 = torch.from_numpy(G["el_img1"][..., :1].copy()).to(dev).requires_grad_(True)
    b1 = torch.from_numpy(G["el_img2"][..., :1].copy()).to(dev).requires_grad_(True)
    loss, delta = events.event_loss(a1, b1, torch.from_numpy(G["el_pols"]).to(dev), use_luma=False, linlog=True, C_thres=-1, event_only=True)
    want, wdelta = eo.event_loss(G["el_img1"][..., :1], G["el_img2"][..., :1], G["el_pols"], 0, 1, -1, 1)
    assert abs(float(loss) - want) <= 2e-4 * abs(want) and np.allclose(delta.cpu().numpy(), wdelta, atol=5e-6)
    loss.backward()
    assert torch.isfinite(a1.grad).all() and float(a1.grad.abs().sum()) > 0


@pytest.mark.gpu
def test_gpu_fused_adam_matches_torch_adam_with_grad_scaler():
    """N4: FusedAdam vs torch.optim.Adam (E-NeRF's hyper-parameters) over 20 steps, driven through a GradScaler, including a
    step with an overflowing gradient that both must skip."""
    from enerf_b200.optim import FusedAdam
    dev = "cuda"
    torch.manual_seed(0)
    shapes = [(100003, 2), (7168,), (5,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.1) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    kw = dict(lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    ref_opt = torch.optim.Adam(ref_p, fused=True, capturable=True, **kw)
    our_opt = FusedAdam(our_p, **kw)
    ref_sc, our_sc = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for it in range(20):
        grads = [torch.randn(s, device=dev) * (10.0 ** ((it % 5) - 3)) for s in shapes]
        if it == 7:
            grads[1][3] = float("inf")
        for ps, opt, sc in ((ref_p, ref_opt, ref_sc), (our_p, our_opt, our_sc)):
            scale = sc.scale(torch.ones((), device=dev))            # also initialises the scaler lazily, as scale(loss) does
            for p, g in zip(ps, grads):
                p.grad = g * scale
            sc.step(opt)
            sc.update()
    assert ref_sc.get_scale() == our_sc.get_scale() == 512.0          # one skipped step halved the scale
    for a, b in zip(ref_p, our_p):
        assert torch.isfinite(b).all()
        err = float((a - b).abs().max())
        assert err <= 2e-6 * float(a.abs().max()) + 1e-8, err
    assert float(our_opt.state[our_p[0]]["step"]) == float(ref_opt.state[ref_p[0]]["step"]) == 19.0


# ------------------------------------------------------------------ N3: event-pair sampler
S = np.load(os.path.join(os.path.dirname(__file__), "golden", "sampler.npz"))


@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_event_pair_sampler_reproduces_reference_collate(case):
    ev = S[case + "_events"]
    eidx, eend, pols, xs, ys = eo.sample_event_pairs(ev, S[case + "_num_succ"], S[case + "_no_succ"], int(S[case + "_acc_max"]),
                                                     S[case + "_u_start"], S[case + "_u_end"])
    assert np.array_equal(pols, S[case + "_pols"][0])
    assert np.all(eend > eidx) and np.all(ev[eidx, 0] == ev[eend, 0]) and np.all(ev[eidx, 1] == ev[eend, 1])      # same pixel, later event
    r = eo.get_event_rays(xs, ys, S[case + "_poses_evs"][eidx][None], S[case + "_poses_evs"][eend][None], S[case + "_intr"])
    for k in ("o1", "d1", "o2", "d2"):
        assert np.allclose(r["rays_evs_" + k], S[case + "_" + k], atol=2e-7), k


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_gpu_event_pair_sampler_matches_reference_collate(case):
    from enerf_b200 import events
    dev = "cuda"
    sampler = events.EventPairSampler(S[case + "_events"], S[case + "_num_succ"], S[case + "_no_succ"], int(S[case + "_acc_max"]),
                                      S[case + "_poses_evs"], device=dev)
    out = sampler.sample(len(S[case + "_u_start"]), S[case + "_intr"], u_start=torch.from_numpy(S[case + "_u_start"]),
                         u_end=torch.from_numpy(S[case + "_u_end"]))
    assert torch.equal(out["pols"].cpu(), torch.from_numpy(S[case + "_pols"]))
    for k in ("o1", "d1", "o2", "d2"):
        assert np.allclose(out["rays_evs_" + k].cpu().numpy(), S[case + "_" + k], atol=3e-7), k
    # free-running draws: valid pairs, right distribution support
    out = sampler.sample(4096)
    ev = S[case + "_events"]
    s, e = out["eidx"].cpu().numpy(), out["eidx_end"].cpu().numpy()
    assert np.all(e > s) and np.all(ev[s, 0] == ev[e, 0]) and np.all(ev[s, 1] == ev[e, 1])
    acc_max = int(S[case + "_acc_max"])
    if acc_max:
        assert (e - s).max() <= acc_max + 1
    assert np.array_equal(out["pols"].cpu().numpy()[0], np.array([ev[a + 1:b + 1, 3].sum() for a, b in zip(s, e)], np.float32))

"""The torch-topology field of nerf/network.py on the tcgen05 kernels: sigma-net 32-64-16 with the trunc_exp
head, colour-net 31(+1)-64-64-C on the masked samples — against torch's own fp16 nn.Linear chain (what the reference runs under
autocast), against a float64 evaluation of the same weights, and at renderer level against the mirror with the tensor-core path
switched off.  Tolerances: 2e-3 relative (one fp16 ulp at the activations' magnitude), gradients 1e-2 .. 3e-2 relative L2."""
import pytest
import torch
import torch.nn.functional as F

from enerf_b200 import field, synthetic
from enerf_b200 import raymarching as rm
from enerf_b200.nerf.network import NeRFNetwork
from tests.gpu_common import DEV, t

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _lin(i, o, seed):
    l = torch.nn.Linear(i, o, bias=False)
    with torch.no_grad():
        l.weight.copy_(torch.randn(o, i, generator=torch.Generator().manual_seed(seed)) * (2.0 / i) ** 0.5)
    return l.to(DEV)


def _round_keep_grad(v):
    """value rounded to fp16, gradient of the unrounded expression (what a fp16 kernel with an exact backward computes)"""
    return v + (v.detach().half().double() - v.detach())


@pytest.mark.parametrize("B", [128, 128 * 301])
def test_density_head_matches_linear_chain(B):
    l0, l1 = _lin(32, 64, 1), _lin(64, 16, 2)
    feat = (torch.randn(B, 32, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3)) * 0.5).half().requires_grad_(True)
    g_sigma = torch.randn(B, device=DEV) * 0.1
    g_h = (torch.randn(B, 16, device=DEV) * 0.1).half()
    g_h[:, 0] = 0
    sigma, h = field.density_head(feat, field.flat_sigma_weights([l0, l1]), 1)
    ((sigma * g_sigma).sum() + (h.float() * g_h.float()).sum()).backward()
    got = (sigma.detach(), h.detach(), feat.grad.clone(), l0.weight.grad.clone(), l1.weight.grad.clone())
    # float64 evaluation of the same network with the same fp16 rounding points (weights, hidden activation, output)
    f64 = feat.detach().double().requires_grad_(True)
    w0, w1 = l0.weight.detach().half().double().requires_grad_(True), l1.weight.detach().half().double().requires_grad_(True)
    a = _round_keep_grad(torch.relu(f64 @ w0.t()))
    hr = _round_keep_grad(a @ w1.t())
    sr = torch.exp(hr[:, 0])
    ((sr * g_sigma.double()).sum() + (hr * g_h.double()).sum()).backward()
    assert float((got[1].double() - hr.detach()).abs().max()) <= 2e-3 * float(hr.detach().abs().max()) + 1e-3
    assert _rel(got[0], sr.detach()) < 2e-3
    assert _rel(got[2], f64.grad) < 1e-2 and _rel(got[3], w0.grad) < 1e-2 and _rel(got[4], w1.grad) < 1e-2
    # torch's own fp16 path (autocast nn.Linear = what the reference executes)
    for l in (l0, l1):
        l.weight.grad = None
    f2 = feat.detach().clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        h2 = l1(F.relu(l0(f2)))
    s2 = torch.exp(h2[:, 0].float())
    ((s2 * g_sigma).sum() + (h2.float() * g_h.float()).sum()).backward()
    assert float((got[1].float() - h2.float()).abs().max()) <= 4e-3 * float(h2.float().abs().max()) + 1e-3
    assert _rel(got[2], f2.grad) < 2e-2 and _rel(got[3], l0.weight.grad) < 2e-2 and _rel(got[4], l1.weight.grad) < 2e-2


@pytest.mark.parametrize("n_ch,frac,dir_div", [(1, 0.3, 1), (3, 0.05, 16), (3, 1.0, 1), (1, 0.0, 1)])
def test_masked_color_matches_module_chain(n_ch, frac, dir_div):
    B = 128 * 40
    torch.manual_seed(n_ch)
    net = NeRFNetwork(bound=1, out_dim_color=n_ch).to(DEV)
    h = (torch.randn(B, 16, device=DEV) * 0.5).half().requires_grad_(True)
    d = torch.randn(B // dir_div, 3, device=DEV)
    d = d / d.norm(dim=-1, keepdim=True)
    mask = torch.rand(B, device=DEV) < frac
    if frac >= 1.0:
        mask[:] = True
    g = torch.randn(B, n_ch, device=DEV)
    x = torch.zeros(B, 3, device=DEV)
    d_view = d[:, None, :].expand(B // dir_div, dir_div, 3) if dir_div > 1 else d
    with torch.autocast("cuda", dtype=torch.float16):
        got = net.color(x, d_view, mask=mask, geo_feat=h[:, 1:], h=h)
    assert got.shape == (B, n_ch)
    if (~mask).any():
        assert float(got[~mask].abs().max()) == 0
    got_g = None
    if mask.any():
        (got * g).sum().backward()
        got_g = (h.grad.clone(), [l.weight.grad.clone() for l in net.color_net])
        for l in net.color_net:
            l.weight.grad = None
    # the reference formulation (nn.Linear, boolean gather / scatter) on the same inputs
    net.use_tensor_cores = False
    h2 = h.detach().clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        want = net.color(x, d_view.reshape(-1, 3), mask=mask, geo_feat=h2[:, 1:])
    assert float((got - want).abs().max()) <= 3e-3
    if got_g is not None:
        (want * g).sum().backward()
        assert _rel(got_g[0], h2.grad) < 2e-2
        for a, l in zip(got_g[1], net.color_net):
            assert _rel(a, l.weight.grad) < 2e-2


@pytest.mark.parametrize("bound,n_ch,upsample", [(1, 3, 0), (3, 1, 0), (2, 1, 32)])
def test_run_on_tensor_cores_matches_linear_formulation(bound, n_ch, upsample):
    """renderer level, fp16 autocast: image / depth / every parameter gradient of the tcgen05 path vs the same mirror on nn.Linear."""
    torch.manual_seed(bound)
    model = NeRFNetwork(bound=bound, out_dim_color=n_ch).to(DEV).train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.5, 0.5)
    if upsample:
        model.eval()
    o, d = synthetic.random_rays(300, bound, seed=9)
    target = torch.rand(300, n_ch, device=DEV)
    res = {}
    for tc in (True, False):
        model.use_tensor_cores = tc
        for p_ in model.parameters():
            p_.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(t(o)[None], t(d)[None], staged=False, num_steps=64, upsample_steps=upsample, bg_color=1, perturb=False,
                               out_dim_color=n_ch)
        loss = ((out["image"][0].float() - target) ** 2).mean()
        (loss * 4096.0).backward()          # what GradScaler does: keeps the fp16 gradients of BOTH formulations out of the subnormal range
        res[tc] = (out["image"][0].detach().float(), out["depth"][0].detach().float(), [p_.grad.clone() for p_ in model.parameters()])
    assert float((res[True][0] - res[False][0]).abs().max()) <= 5e-3
    assert float((res[True][1] - res[False][1]).abs().max()) <= 5e-3
    for a, b in zip(res[True][2], res[False][2]):
        assert _rel(a, b) < 3e-2, _rel(a, b)
    assert float(res[True][2][0].abs().sum()) > 0


def test_weighted_sum_and_row_moves():
    from enerf_b200 import _lib
    N, T = 257, 70
    for C in (1, 3, 4):
        w = torch.rand(N, T, device=DEV, requires_grad=True)
        rgb = torch.rand(N, T, C, device=DEV, requires_grad=True)
        g = torch.randn(N, C, device=DEV)
        img = rm.weighted_sum(w, rgb)
        (img * g).sum().backward()
        w2, r2 = w.detach().clone().requires_grad_(True), rgb.detach().clone().requires_grad_(True)
        ref = (w2.unsqueeze(-1) * r2).sum(-2)
        (ref * g).sum().backward()
        assert torch.allclose(img, ref, atol=1e-5, rtol=1e-5)
        assert torch.allclose(w.grad, w2.grad, atol=1e-5, rtol=1e-5) and torch.allclose(rgb.grad, r2.grad, atol=1e-6, rtol=1e-5)
    src = torch.randn(1000, 32, device=DEV).half()
    idx = torch.randperm(1000, device=DEV)[:333].sort().values.int()
    dst = torch.full((384, 32), 9.0, device=DEV, dtype=torch.half)
    _lib.call("enerf_gather_rows", _lib.ptr(src), _lib.ptr(idx), 333, 384, 64, _lib.ptr(dst), None, _lib.stream())
    assert torch.equal(dst[:333], src[idx.long()]) and float(dst[333:].abs().max()) == 0
    back = torch.zeros(1000, 32, device=DEV, dtype=torch.half)
    _lib.call("enerf_scatter_rows", _lib.ptr(dst), _lib.ptr(idx), 333, 64, _lib.ptr(back), None, _lib.stream())
    # the same with the row count on the device and capacities on the host side
    cnt = torch.tensor([200], dtype=torch.int32, device=DEV)
    dst2 = torch.full((384, 32), 9.0, device=DEV, dtype=torch.half)
    _lib.call("enerf_gather_rows", _lib.ptr(src), _lib.ptr(idx), 333, 384, 64, _lib.ptr(dst2), _lib.ptr(cnt), _lib.stream())
    assert torch.equal(dst2[:200], src[idx.long()][:200]) and float(dst2[200:256].abs().max()) == 0 and float(dst2[256:].float().min()) == 9.0
    assert torch.equal(back[idx.long()], src[idx.long()])
    rest = torch.ones(1000, dtype=torch.bool, device=DEV)
    rest[idx.long()] = False
    assert float(back[rest].abs().max()) == 0


def test_grid_density_paths_agree():
    """the density-only kernel used by update_extra_state vs `density()['sigma']`, both topologies"""
    from enerf_b200.nerf.network_ff import NeRFNetwork as FFNet
    x = (torch.rand(128 * 50, 3, device=DEV) * 2 - 1)
    for cls in (NeRFNetwork, FFNet):
        torch.manual_seed(1)
        m = cls(bound=1, cuda_ray=True, out_dim_color=1).to(DEV).train()
        with torch.no_grad():
            m.encoder.embeddings.uniform_(-0.5, 0.5)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            a = m._grid_density(x)
            b = m.density(x)['sigma'].float()
        assert a.dtype == torch.float32 and float((a - b).abs().max()) <= 2e-3 * float(b.abs().max())


@pytest.mark.parametrize("n_rays,T", [(100, 33), (1, 7), (257, 128)])
def test_run_on_tensor_cores_ragged_sizes(n_rays, T):
    """sample counts that are not a multiple of the 128-row MLP tile (padding inside density(), capacity padding in the colour batch)"""
    torch.manual_seed(7)
    model = NeRFNetwork(bound=1, out_dim_color=1).to(DEV).train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.5, 0.5)
    o, d = synthetic.random_rays(n_rays, 1, seed=3)
    res = {}
    for tc in (True, False):
        model.use_tensor_cores = tc
        for p_ in model.parameters():
            p_.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(t(o)[None], t(d)[None], staged=False, num_steps=T, upsample_steps=0, bg_color=1, perturb=False, out_dim_color=1)
        (out["image"].float().sum() * 256.0).backward()
        res[tc] = (out["image"][0].detach().float(), [p_.grad.clone() for p_ in model.parameters()])
    assert res[True][0].shape == (n_rays, 1)
    assert float((res[True][0] - res[False][0]).abs().max()) <= 5e-3
    for a, b in zip(res[True][1], res[False][1]):
        assert torch.isfinite(a).all() and _rel(a, b) < 5e-2, _rel(a, b)


@pytest.mark.parametrize("n_ch", [1, 3])
def test_forward_all_samples_matches_linear_formulation(n_ch):
    """`forward(x, d)` (what run_cuda calls for `cuda_ray` without `ff`): colour-net on every row, ragged row count"""
    torch.manual_seed(11)
    model = NeRFNetwork(bound=1, out_dim_color=n_ch).to(DEV).train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.5, 0.5)
    B = 128 * 9 + 37
    x = torch.rand(B, 3, device=DEV) * 2 - 1
    d = torch.randn(B, 3, device=DEV)
    d = d / d.norm(dim=-1, keepdim=True)
    gs, gr = torch.randn(B, device=DEV), torch.randn(B, n_ch, device=DEV)
    res = {}
    for tc in (True, False):
        model.use_tensor_cores = tc
        for p_ in model.parameters():
            p_.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            sigma, rgb = model(x, d)
        ((sigma.float() * gs).sum() + (rgb.float() * gr).sum()).backward()
        res[tc] = (sigma.detach().float(), rgb.detach().float(), [p_.grad.clone() for p_ in model.parameters()])
    assert res[True][1].shape == (B, n_ch)
    assert _rel(res[True][0], res[False][0]) < 3e-3 and float((res[True][1] - res[False][1]).abs().max()) <= 3e-3
    for a, b in zip(res[True][2], res[False][2]):
        assert _rel(a, b) < 3e-2, _rel(a, b)


def test_cuda_ray_training_step_on_tensor_cores():
    """`--cuda_ray` without `--ff`: run_cuda's training branch queries forward(x, d) of the nerf/network.py topology on every marched
    sample; tensor-core path vs the nn.Linear formulation, image and every parameter gradient"""
    torch.manual_seed(5)
    model = NeRFNetwork(bound=1, cuda_ray=True, out_dim_color=3, density_thresh=0.01).to(DEV).train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.5, 0.5)
    torch.manual_seed(2)
    with torch.autocast("cuda", dtype=torch.float16):
        model.update_extra_state()
    o, d = synthetic.random_rays(300, 1, seed=9)
    res = {}
    for tc in (True, False):
        model.use_tensor_cores = tc
        for p_ in model.parameters():
            p_.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(t(o)[None], t(d)[None], staged=False, bg_color=1, perturb=True, force_all_rays=True, out_dim_color=3)
        (out["image"].float().sum() * 64.0).backward()
        res[tc] = (out["image"][0].detach().float(), out["depth"][0].detach().float(), [p_.grad.clone() for p_ in model.parameters()])
    assert int(model.step_counter[(model.local_step - 1) % 16, 0]) > 1000          # the rays hit occupied cells
    assert float((res[True][0] - res[False][0]).abs().max()) <= 5e-3
    assert float((res[True][1] - res[False][1]).abs().max()) <= 1e-5               # same samples
    for a, b in zip(res[True][2], res[False][2]):
        assert torch.isfinite(a).all() and _rel(a, b) < 5e-2, _rel(a, b)

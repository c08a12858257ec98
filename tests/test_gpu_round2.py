"""GPU tests added in round 2 for the advisor's findings: the fp16 table shadow under FusedAdam (eager and
CUDA-graph replay) and the last-sample delta of the fused run() integrator after PDF upsampling."""
import numpy as np
import pytest
import torch

from enerf_b200 import raymarching as rm
from enerf_b200.gridencoder import GridEncoder
from enerf_b200.gridencoder.grid import half_shadow
from enerf_b200.optim import FusedAdam
from tests.gpu_common import DEV, n

pytestmark = pytest.mark.gpu


def _encoder_step_fn(enc, opt, scaler, x, target):
    def step():
        with torch.autocast("cuda", dtype=torch.float16):
            f = enc(x, bound=1)
        loss = ((f.float().sum(-1) - target) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss
    return step


@pytest.mark.parametrize("graphed", [False, True])
def test_fused_adam_keeps_fp16_table_current(graphed):
    """FusedAdam updates the table through its raw pointer, which does not bump `_version`; the fp16 copy the
    kernels read must follow anyway.  After every step the shadow equals half(parameter), the encoder output moves, the loss drops;
    a `copy_` into the parameter (EMA swap, checkpoint load) is picked up by the next forward."""
    torch.manual_seed(0)
    enc = GridEncoder(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=15, desired_resolution=512).to(DEV)
    with torch.no_grad():
        enc.embeddings.uniform_(-0.1, 0.1)
    x = torch.rand(4096, 3, device=DEV) * 2 - 1
    target = torch.sin(x.sum(-1) * 3)
    opt = FusedAdam(enc.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda", init_scale=128.0)
    step = _encoder_step_fn(enc, opt, scaler, x, target)
    losses = [float(step())]
    assert half_shadow(enc.embeddings) is not None
    if graphed:
        from enerf_b200.graphs import GraphedStep
        g = GraphedStep(lambda: step(), [], warmup=2)
        run = lambda: g()                                       # noqa: E731
    else:
        run = step
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        before = enc(x, bound=1).clone()
    for _ in range(6):
        losses.append(float(run()))
        torch.cuda.synchronize()
        sh = half_shadow(enc.embeddings)[0]
        assert torch.equal(sh, enc.embeddings.detach().half()), "fp16 shadow is stale after a FusedAdam step"
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        after = enc(x, bound=1)
    assert not torch.equal(before, after), "the encoder does not see its own updates"
    assert min(losses[1:]) < 0.5 * losses[0], losses
    # a writer that goes through torch (version bump) invalidates the shadow: the next forward re-casts
    with torch.no_grad():
        enc.embeddings.copy_(torch.zeros_like(enc.embeddings))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        zeroed = enc(x, bound=1)
    assert float(zeroed.abs().max()) == 0.0


def test_two_encoders_keep_separate_shadows():
    """one slot per parameter — a second encoder (encoder_bg) must not evict the first one's copy."""
    a = GridEncoder(num_levels=4, log2_hashmap_size=10, desired_resolution=64).to(DEV)
    b = GridEncoder(num_levels=4, log2_hashmap_size=10, desired_resolution=64).to(DEV)
    x = torch.rand(256, 3, device=DEV) * 2 - 1
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        a(x), b(x)
        pa, pb = half_shadow(a.embeddings)[0].data_ptr(), half_shadow(b.embeddings)[0].data_ptr()
        a(x), b(x)
    assert pa != pb and half_shadow(a.embeddings)[0].data_ptr() == pa and half_shadow(b.embeddings)[0].data_ptr() == pb


def _weights_reference(sig, z, nears, fars, ds, num_steps):
    """renderer.py:230-234 with the reference's last delta = (far-near)/num_steps (the COARSE count, renderer.py:177,231)"""
    sd = (fars - nears) / num_steps
    deltas = torch.cat([z[:, 1:] - z[:, :-1], sd[:, None]], -1)
    alphas = 1 - torch.exp(-deltas * ds * sig)
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-15], -1)
    w = alphas * torch.cumprod(shifted, -1)[:, :-1]
    depth = (w * ((z - nears[:, None]) / (fars - nears)[:, None]).clamp(0, 1)).sum(-1)
    return w, w.sum(-1), depth


def test_composite_uniform_last_delta_after_upsampling():
    """with upsample_steps > 0 the rows are num_steps + upsample_steps long but the last sample's delta stays
    (far-near)/num_steps.  Forward and gradient vs the float64 formula; also shows the old behaviour (T_dist = T) is different."""
    N, T0, Tu = 64, 48, 32
    g = torch.Generator().manual_seed(3)
    nears = torch.rand(N, generator=g).double() + 0.2
    fars = nears + 1 + torch.rand(N, generator=g).double() * 3
    z = torch.sort(nears[:, None] + (fars - nears)[:, None] * torch.rand(N, T0 + Tu, generator=g).double(), dim=1).values
    sig = (torch.rand(N, T0 + Tu, generator=g) * 0.3).double()  # nearly transparent up to the last sample ...
    sig[:, -1] = 40.0                                           # ... so that its delta decides a visible weight
    sig_ref = sig.clone().requires_grad_(True)
    w, ws, depth = _weights_reference(sig_ref, z, nears, fars, 1.0, T0)
    gw = torch.randn(N, T0 + Tu, generator=g).double()
    ((w * gw).sum() + ws.sum() + depth.sum()).backward()

    s_gpu = sig.float().to(DEV).requires_grad_(True)
    args = (z.float().to(DEV), nears.float().to(DEV), fars.float().to(DEV), 1.0)
    W, WS, D = rm.composite_uniform(s_gpu, *args, T0)
    assert np.allclose(n(W), w.detach().numpy(), atol=2e-6, rtol=1e-4)
    assert np.allclose(n(WS), ws.detach().numpy(), atol=1e-5) and np.allclose(n(D), depth.detach().numpy(), atol=1e-5)
    ((W * gw.float().to(DEV)).sum() + WS.sum() + D.sum()).backward()
    ref_g = sig_ref.grad.numpy()
    assert np.allclose(n(s_gpu.grad), ref_g, atol=2e-5 * max(1.0, np.abs(ref_g).max()), rtol=2e-3)
    W_old, _, _ = rm.composite_uniform(s_gpu.detach(), *args)    # num_steps = 0 -> (far-near)/T
    assert float((W_old[:, -1] - W[:, -1]).abs().max()) > 1e-3


@pytest.mark.parametrize("n_step", [32, 33, 100, 257])
def test_warp_per_ray_inference_compositor_matches_oracle(n_step):
    """composite_rays with n_step >= 32 runs one warp per ray (the mirror's inference loop marches tens to hundreds of steps
    per round, for which the thread-per-ray kernel reads strided).  Same results as the sequential oracle (raymarching.cu:842-899),
    including both stopping rules (padding with delta == 0, transmittance < 1e-5), the rays_t = -1 marking and a device-side alive count
    smaller than the launch bound."""
    from oracle import oracle
    from tests.gpu_common import scene, t
    bound = 2
    sc = scene(600, bound, seed=7)
    N, n_alive = 600, 450
    rng = np.random.default_rng(n_step)
    alive_np = rng.permutation(N)[:n_alive].astype(np.int32)
    t_np = sc["nears"][alive_np].copy()
    wx, wd, wdl = oracle.march_rays(n_alive, n_step, alive_np, t_np, sc["o"], sc["d"], bound, sc["bits"], sc["cascade"], 128, sc["nears"], sc["fars"], perturb=0)
    m = n_alive * n_step
    # densities: a third of the rays nearly transparent, a third terminating by the transmittance rule, a third mixed
    kind = rng.integers(0, 3, n_alive)
    sig = (rng.uniform(0, 1, (n_alive, n_step)) * np.array([0.5, 3000.0, 150.0])[kind][:, None]).astype(np.float32).reshape(-1)
    rgb = rng.random((m, 3)).astype(np.float32)
    ws0, d0, im0 = rng.random(N).astype(np.float32) * 0.3, rng.random(N).astype(np.float32), rng.random((N, 3)).astype(np.float32)
    wt, wws, wdp, wim = oracle.composite_rays(n_alive, n_step, alive_np, t_np, sig, rgb, wdl, ws0, d0, im0)
    assert (wt < 0).any() and (wt >= 0).any()
    for bound_n, dev_count in ((n_alive, None), (N, n_alive)):
        gt, gws_, gdp, gim = torch.zeros(N, device=DEV), t(ws0), t(d0), t(im0)
        gt[:n_alive] = t(t_np)
        alive_pad = torch.zeros(N, dtype=torch.int32, device=DEV)
        alive_pad[:n_alive] = t(alive_np)
        pad = bound_n * n_step - m
        sig_g = torch.cat([t(sig), torch.zeros(pad, device=DEV)])
        rgb_g = torch.cat([t(rgb), torch.zeros(pad, 3, device=DEV)])
        dl_g = torch.cat([t(wdl), torch.zeros(pad, 2, device=DEV)])
        cnt = None if dev_count is None else torch.tensor([dev_count], dtype=torch.int32, device=DEV)
        rm.composite_rays(bound_n, n_step, alive_pad, gt, sig_g, rgb_g, dl_g, gws_, gdp, gim, cnt)
        got_t = n(gt)[:n_alive]
        assert np.array_equal(got_t < 0, wt < 0)
        assert np.allclose(got_t, wt, atol=1e-5) and np.allclose(n(gws_), wws, atol=2e-6) and np.allclose(n(gdp), wdp, atol=2e-5)
        assert np.allclose(n(gim), wim, atol=2e-6)


def test_inference_loop_without_per_round_sync_renders_the_same_image():
    """run_cuda inference: reference policy (n_step <= 8, counter read after every compaction) vs large rounds with the alive count kept
    on the device and read every 4 rounds — same image and depth (perturb off), far fewer host reads."""
    from enerf_b200 import synthetic
    from enerf_b200.nerf.network_ff import NeRFNetwork
    from tests.gpu_common import t
    bound = 2
    torch.manual_seed(0)
    model = NeRFNetwork(bound=bound, cuda_ray=True, out_dim_color=3).to(DEV).eval()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-2.0, 2.0)           # dense enough for rays to terminate at different depths
    grid = synthetic.ball_density_grid(bound, model.cascade)
    model.density_grid.copy_(t(grid))
    model.density_bitfield.copy_(t(synthetic.packbits_np(grid)))
    o, d = synthetic.random_rays(3000, bound, seed=4)
    res = {}
    for name, batch, every in (("reference_policy", 0, 1), ("large_rounds", 1 << 17, 1), ("device_count", 1 << 17, 4)):
        model.inference_batch_samples, model.inference_sync_every = batch, every
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out = model.render(t(o)[None], t(d)[None], staged=False, bg_color=1, perturb=False, out_dim_color=3)
        res[name] = (out["image"][0].float(), out["depth"][0].float(), dict(model.last_render_stats))
    a = res["reference_policy"]
    for name in ("large_rounds", "device_count"):
        b = res[name]
        assert float((a[0] - b[0]).abs().max()) <= 2e-3 and float((a[1] - b[1]).abs().max()) <= 2e-3, name
        assert b[2]["samples"] >= a[2]["samples"], (a[2], b[2])           # a ray that dies mid-round still owns its slot until the round ends
    assert res["device_count"][2]["host_syncs"] * 3 <= res["large_rounds"][2]["host_syncs"]
    assert res["large_rounds"][2]["iterations"] * 4 <= a[2]["iterations"]

"""The fused inference field (csrc/field_infer.cu: hash-grid gather + sigma-net + colour-net as one kernel) against the chain it
replaces — grid_encode_forward -> field_sigma_forward -> field_color_forward, each parity-tested on its own against the oracle and
the reference build (test_gpu_encoders.py, test_gpu_ffmlp.py).  Same arithmetic, same rounding points: the outputs are compared bit
for bit, on ragged sizes, out-of-range positions, every channel count, both grid types, and on a whole rendered frame."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _model(bound=3, n_ch=1, gridtype="hash", seed=0):
    from enerf_b200.gridencoder import GridEncoder
    from enerf_b200.nerf.network_ff import NeRFNetwork
    torch.manual_seed(seed)
    m = NeRFNetwork(encoding="hashgrid", bound=bound, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1,
                    out_dim_color=n_ch).to(DEV)
    if gridtype != "hash":
        m.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                                desired_resolution=2048 * bound, gridtype=gridtype).to(DEV)
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-0.5, 0.5)          # "trained-like": features of order 1, so the nets' outputs spread out
    m.eval()
    return m


def _samples(n, bound, seed, frac_outside=0.05):
    g = torch.Generator(device="cpu").manual_seed(seed)
    # ray-like runs (consecutive samples close together, as the marcher emits them) plus some positions outside the box
    starts = (torch.rand(n, 3, generator=g) * 2 - 1) * bound
    x = starts.clone()
    run = 64
    for i in range(0, n, run):
        k = min(run, n - i)
        step = torch.randn(3, generator=g) * 0.004 * bound
        x[i:i + k] = starts[i] + torch.arange(k).unsqueeze(1) * step
    x = x.clamp(-bound, bound)
    outside = torch.rand(n, generator=g) < frac_outside
    x[outside] *= 1.5                                      # beyond [-bound, bound] in some coordinate: zero features (gridencoder.cu:250-256)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=1, keepdim=True)
    return x.float().to(DEV).contiguous(), d.float().to(DEV).contiguous()


def _both(m, x, d):
    """(fused kernel, unfused chain on the rows padded to a multiple of 128)"""
    n = x.shape[0]
    pad = (-n) % 128
    xp = torch.cat([x, x.new_zeros(pad, 3)]) if pad else x
    dp = torch.cat([d, d.new_zeros(pad, 3)]) if pad else d
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        m.fuse_infer = True
        s1, c1 = m(x, d)
        m.fuse_infer = False
        s0, c0 = m(xp, dp)
        m.fuse_infer = True
    return (s1, c1), (s0[:n], c0[:n])


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000, 148 * 128 * 3 + 77, 1 << 20])
def test_fused_field_is_the_unfused_chain_bit_for_bit(n):
    from enerf_b200 import _lib
    m = _model(bound=3, n_ch=1)
    x, d = _samples(n, 3, seed=n)
    launches = _lib.launch_count()
    (s1, c1), (s0, c0) = _both(m, x, d)
    assert _lib.launch_count() - launches >= 4            # 1 fused + 3 unfused kernels at least: both paths ran native code
    assert s1.shape == (n,) and c1.shape == (n, 1) and s1.dtype == torch.float32 and c1.dtype == torch.float32
    assert torch.equal(s1, s0)
    assert torch.equal(c1, c0)
    assert bool(torch.isfinite(s1).all()) and float(c1.min()) >= 0.0 and float(c1.max()) <= 1.0
    if n > 1:
        assert float(s1.std()) > 0 and float(c1.std()) > 0     # not a degenerate comparison


@pytest.mark.parametrize("n_ch", [1, 3, 4])
@pytest.mark.parametrize("bound", [1, 2])
def test_channel_counts_and_bounds(n_ch, bound):
    m = _model(bound=bound, n_ch=n_ch, seed=n_ch)
    x, d = _samples(70000 + 13 * n_ch, bound, seed=100 + n_ch)
    (s1, c1), (s0, c0) = _both(m, x, d)
    assert torch.equal(s1, s0) and torch.equal(c1, c0)


def test_tiled_grid_and_unit_bound():
    m = _model(bound=1, n_ch=3, gridtype="tiled")
    x, d = _samples(33333, 1, seed=5)
    (s1, c1), (s0, c0) = _both(m, x, d)
    assert torch.equal(s1, s0) and torch.equal(c1, c0)


def test_outputs_follow_the_table_after_an_optimizer_step():
    """the kernel reads the fp16 shadow FusedAdam keeps current: a step changes what the fused field returns"""
    from enerf_b200.optim import FusedAdam
    m = _model(bound=1, n_ch=1)
    x, d = _samples(4096, 1, seed=9, frac_outside=0.0)
    (s_before, _), _ = _both(m, x, d)
    m.train()
    opt = FusedAdam(m.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    with torch.autocast("cuda", dtype=torch.float16):
        sig, rgb = m(x, d)
        (sig.mean() + rgb.mean()).backward()
    opt.step()
    m.eval()
    (s1, c1), (s0, c0) = _both(m, x, d)
    assert torch.equal(s1, s0) and torch.equal(c1, c0)
    assert not torch.equal(s1, s_before)


def test_refused_shapes_and_fallback():
    from enerf_b200 import _lib
    from enerf_b200._lib import ptr, stream
    m = _model(bound=1, n_ch=1)
    x, d = _samples(256, 1, seed=1)
    out = torch.empty(256, device=DEV)
    tab = m.encoder.embeddings.detach().half()
    w = torch.zeros(64 * (32 + 64 * 3 + 16), dtype=torch.float16, device=DEV)
    with pytest.raises(RuntimeError, match="16 levels"):
        _lib.call("enerf_field_infer", ptr(x), 1.0, 0.5, ptr(d), ptr(tab), ptr(m.encoder.offsets), 8, 2, 0.5, 16, 0, ptr(w), 2, ptr(w), 3, 256, 1,
                  ptr(out), ptr(out), stream())
    with pytest.raises(RuntimeError, match="FFMLP"):
        _lib.call("enerf_field_infer", ptr(x), 1.0, 0.5, ptr(d), ptr(tab), ptr(m.encoder.offsets), 16, 2, 0.5, 16, 0, ptr(w), 3, ptr(w), 3, 256, 1,
                  ptr(out), ptr(out), stream())
    # an empty batch is a no-op
    _lib.call("enerf_field_infer", ptr(x), 1.0, 0.5, ptr(d), ptr(tab), ptr(m.encoder.offsets), 16, 2, 0.5, 16, 0, ptr(w), 2, ptr(w), 3, 0, 1,
              ptr(out), ptr(out), stream())
    # with gradients enabled the module keeps the differentiable chain
    m.train()
    with torch.autocast("cuda", dtype=torch.float16):
        sig, _ = m(x, d)
    assert sig.requires_grad


def test_rendered_frame_is_identical():
    """run_cuda's inference loop (renderer.py:344-401) through the fused field and through the chain: same image, same depth"""
    from enerf_b200 import synthetic
    bound = 2
    m = _model(bound=bound, n_ch=3)
    grid = synthetic.ball_density_grid(bound, m.cascade)
    m.density_grid.copy_(torch.from_numpy(grid))
    m.density_bitfield.copy_(torch.from_numpy(synthetic.packbits_np(grid)))
    pose = synthetic.look_at_poses(1, 0.6 * bound, seed=3)[0]
    o_np, d_np = synthetic.pinhole_rays(pose, 96, 96, 50.0, np.arange(96 * 96))
    o, d = torch.from_numpy(o_np).to(DEV), torch.from_numpy(d_np).to(DEV)
    outs = []
    for fuse in (True, False):
        m.fuse_infer = fuse
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs.append(m.render(o.unsqueeze(0), d.unsqueeze(0), staged=False, bg_color=1, perturb=False, dt_gamma=0, max_steps=256))
    assert torch.equal(outs[0]["image"], outs[1]["image"])
    assert torch.equal(outs[0]["depth"], outs[1]["depth"])
    assert float(outs[0]["image"].std()) > 0


@pytest.mark.parametrize("live_units", [0, 1, 37, 500, 10 ** 6])
def test_device_side_row_count_limits_the_rows_evaluated(live_units):
    """enerf_field_infer_alive: only the first min(S, *n_units_dev * rows_per_unit) rows are evaluated and written — the bits of the
    full call — and nothing behind them is touched (the inference loop sizes a round by a stale, larger alive count)"""
    from enerf_b200 import field
    m = _model(bound=3, n_ch=3)
    per_unit, units = 26, 500
    S = units * per_unit                                   # 13 000 rows = 102 tiles: fewer tiles than CTAs once the count drops
    x, d = _samples(S, 3, seed=11)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        s_full, c_full = field.fused_infer(x, d, m.encoder, m.bound, m.sigma_net.weights, m.color_net.weights, 3)
        count = torch.tensor([live_units], dtype=torch.int32, device=DEV)
        s_lim = torch.full((S,), -7.0, dtype=torch.float32, device=DEV)
        c_lim = torch.full((S, 3), -7.0, dtype=torch.float32, device=DEV)
        field.fused_infer(x, d, m.encoder, m.bound, m.sigma_net.weights, m.color_net.weights, 3, alive=(count, per_unit), out=(s_lim, c_lim))
    torch.cuda.synchronize()
    live = min(S, live_units * per_unit)
    assert float(s_full.min()) >= 0.0 and float(c_full.min()) >= 0.0           # no result looks like the sentinel
    assert torch.equal(s_lim[:live], s_full[:live]) and torch.equal(c_lim[:live], c_full[:live])
    assert bool((s_lim[live:] == -7.0).all()) and bool((c_lim[live:] == -7.0).all())
    with pytest.raises(ValueError):
        field.fused_infer(x, d, m.encoder, m.bound, m.sigma_net.weights, m.color_net.weights, 3, alive=(count.long(), per_unit))

"""GPU parity: hash-grid and SH encoders vs the CPU oracle and the reference's CUDA build.

Tolerances.  Hash grid forward: the oracle is fed the device's per-level scales, after which the
operation sequence is identical -> bit-exact expected; asserted as max |diff| <= 1 ulp of the
output type with >= 99.9 % of elements identical.  Backward: sums of up to thousands of atomics
in arbitrary order -> compared with the exact (float64) sum, 1e-4 relative to the largest
gradient for fp32 accumulation, 2e-2 for the reference-style fp16 atomics.  SH: fp32 within
2e-6 of the closed forms; fp16 outputs within 1 fp16 ulp of the fp32 value (the reference
evaluates in fp16 and is up to ~4 ulp away from it)."""
import numpy as np
import pytest
import torch

from enerf_b200 import gridencoder, shencoder
from enerf_b200.backends import gridencoder_backend as GB
from enerf_b200.backends import shencoder_backend as SB
from oracle import oracle
from tests.gpu_common import DEV, gpu_level_scales, n, ref_mod, scene, t

pytestmark = pytest.mark.gpu


def _table(bound, dtype, seed=0, L=16, log2T=19, D=3, C=2, scale=1.0):
    pls = oracle.per_level_scale_for(2048 * bound, 16, L) if L > 1 else 2.0
    offsets = oracle.grid_offsets(D, L, pls, 16, log2T)
    rng = np.random.default_rng(seed)
    emb = (rng.uniform(-1, 1, size=(offsets[-1], C)) * scale).astype(dtype)
    return pls, offsets, emb


def _marched_points(bound, n_rays=256, seed=0):
    sc = scene(n_rays, bound, seed)
    xyzs, _, _, rays, counter = oracle.march_rays_train(sc["o"], sc["d"], bound, sc["bits"], sc["cascade"], 128, sc["nears"], sc["fars"], perturb=True)
    x = xyzs[:counter[0]]
    return ((x + bound) / (2 * bound)).astype(np.float32)


def _fwd(x, emb, offsets, pls, layout, cg=False, gridtype=0):
    B, D = x.shape
    L = len(offsets) - 1
    C = emb.shape[1]
    te = t(emb)
    out = torch.empty((L, B, C) if layout == 0 else (B, L * C), device=DEV, dtype=te.dtype)
    dy = torch.empty(B, L * D * C, device=DEV, dtype=te.dtype) if cg else torch.empty(1, device=DEV, dtype=te.dtype)
    GB.grid_encode_forward(t(x), te, t(offsets), out, B, D, C, L, np.log2(pls), 16, cg, dy, gridtype, layout)
    return out, dy


@pytest.mark.parametrize("bound", [1, 2, 3])
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_grid_forward_matches_oracle(bound, dtype):
    pls, offsets, emb = _table(bound, dtype, seed=bound)
    rng = np.random.default_rng(bound)
    x = np.concatenate([_marched_points(bound)[:20000], rng.random((4097, 3)).astype(np.float32),
                        np.array([[0, 0, 0], [1, 1, 1], [0.5, 1.0, 0.0], [1.0001, 0.5, 0.5], [0.5, -1e-7, 0.5]], np.float32)])
    scales = gpu_level_scales(pls, 16, 16)
    want, _ = oracle.grid_encode_forward(x, emb, offsets, pls, 16, level_scales=scales)
    got0, _ = _fwd(x, emb, offsets, pls, 0)
    got1, _ = _fwd(x, emb, offsets, pls, 1)
    g0 = n(got0)
    assert np.array_equal(n(got1).reshape(len(x), 16, 2).transpose(1, 0, 2), g0)        # the two layouts agree
    diff = np.abs(g0.astype(np.float64) - want.astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(dtype)).astype(np.float64)
    same = (g0 == want).mean()
    assert same >= 0.999, f"only {same:.5f} identical"
    assert np.all(diff <= ulp), f"max diff {diff.max()} ({(diff / ulp).max():.1f} ulp)"
    assert np.all(g0[:, -2:] == 0)                                                      # out-of-range rows are zero


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_grid_forward_matches_reference_build(dtype):
    R = ref_mod("_gridencoder")
    if R is None:
        pytest.skip("oracle/_ref not built")
    for bound in (1, 3):
        pls, offsets, emb = _table(bound, dtype, seed=10 + bound)
        x = np.concatenate([_marched_points(bound, 512, seed=1), np.random.default_rng(2).random((50000, 3)).astype(np.float32)])
        B = len(x)
        te, tx, to = t(emb), t(x), t(offsets)
        rout = torch.empty(16, B, 2, device=DEV, dtype=te.dtype)
        rdy = torch.empty(B, 16 * 3 * 2, device=DEV, dtype=te.dtype)
        R.grid_encode_forward(tx, te, to, rout, B, 3, 2, 16, float(np.log2(pls)), 16, True, rdy, 0)
        gout, gdy = _fwd(x, emb, offsets, pls, 0, cg=True)
        assert torch.equal(rout, gout), f"forward differs from reference: max {float((rout.float() - gout.float()).abs().max())}"
        assert torch.equal(rdy, gdy), f"dy_dx differs from reference: max {float((rdy.float() - gdy.float()).abs().max())}"


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_grid_forward_hoisted_kernel_is_bit_identical_to_generic(dtype):
    """The two hot-path kernels — k_grid_fwd_w (mode 1: a warp walks all levels of its 32 samples) and k_grid_fwd3 (mode 2: one warp per
    (32 samples, level)) — vs the generic k_grid_fwd (mode 0), and all vs the reference build when it is available."""
    from enerf_b200 import _lib
    MODES = (1, 2)
    R = ref_mod("_gridencoder")
    rng = np.random.default_rng(12)
    for bound, C, L, gridtype, log2T in [(3, 2, 16, 0, 19), (1, 2, 16, 0, 19), (2, 2, 16, 1, 19), (1, 4, 8, 0, 14), (2, 1, 16, 0, 12), (1, 8, 4, 1, 9)]:
        pls = oracle.per_level_scale_for(2048 * bound, 16, L)
        offsets = oracle.grid_offsets(3, L, pls, 16, log2T)
        emb = rng.uniform(-1, 1, (offsets[-1], C)).astype(dtype)
        x = np.concatenate([_marched_points(bound, 96, seed=3)[:30000], rng.random((10001, 3)).astype(np.float32),
                            np.array([[0, 0, 0], [1, 1, 1], [1.0001, 0.5, 0.5]], np.float32)])
        B = len(x)
        outs = {}
        try:
            for mode in MODES + (0,):
                _lib.call("enerf_grid_set_forward_mode", mode)
                for layout in (0, 1):
                    o, _ = _fwd(x, emb, offsets, pls, layout, gridtype=gridtype)
                    outs[(mode, layout)] = o
        finally:
            _lib.call("enerf_grid_set_forward_mode", 1)
        for mode in MODES:
            assert torch.equal(outs[(mode, 0)], outs[(0, 0)]), (mode, bound, C, L, gridtype)
            assert torch.equal(outs[(mode, 1)], outs[(0, 1)]), (mode, bound, C, L, gridtype)
        assert torch.equal(outs[(1, 1)].view(B, L, C).permute(1, 0, 2), outs[(1, 0)])
        if R is not None:
            te = t(emb)
            rout = torch.empty(L, B, C, device=DEV, dtype=te.dtype)
            R.grid_encode_forward(t(x), te, t(offsets), rout, B, 3, C, L, float(np.log2(pls)), 16, False, torch.empty(1, device=DEV, dtype=te.dtype), gridtype)
            assert torch.equal(rout, outs[(1, 0)]), f"hoisted kernel differs from the reference build {(bound, C, L, gridtype)}"


def test_grid_forward_generic_shapes():
    # D=2 (background encoder), C in {1,4,8}, L=4, tiled grid, small tables
    rng = np.random.default_rng(5)
    for D, C, L, gridtype, log2T in [(2, 2, 4, 0, 19), (3, 1, 5, 0, 12), (3, 4, 6, 0, 14), (3, 8, 3, 1, 10), (2, 4, 7, 1, 8)]:
        pls = 1.7
        offsets = oracle.grid_offsets(D, L, pls, 16, log2T)
        emb = rng.uniform(-1, 1, (offsets[-1], C)).astype(np.float32)
        x = rng.random((1531, D)).astype(np.float32)
        want, wdy = oracle.grid_encode_forward(x, emb, offsets, pls, 16, calc_grad_inputs=True, gridtype=gridtype,
                                               level_scales=gpu_level_scales(pls, 16, L))
        B = len(x)
        out = torch.empty(L, B, C, device=DEV)
        dy = torch.empty(B, L * D * C, device=DEV)
        GB.grid_encode_forward(t(x), t(emb), t(offsets), out, B, D, C, L, np.log2(pls), 16, True, dy, gridtype, 0)
        assert np.allclose(n(out), want, atol=1e-6), (D, C, L, gridtype)
        assert np.allclose(n(dy).reshape(B, L, D, C), wdy, atol=1e-3, rtol=1e-4), (D, C, L, gridtype)


@pytest.fixture(params=[1, 0], ids=["walk", "percorner"])
def scatter_mode(request):
    from enerf_b200 import _lib
    _lib.call("enerf_grid_set_backward_mode", request.param)
    yield request.param
    _lib.call("enerf_grid_set_backward_mode", 1)


@pytest.mark.parametrize("dtype,grad_dtype,tol", [(np.float32, torch.float32, 1e-4), (np.float16, torch.float32, 1e-4), (np.float16, torch.float16, 2e-2)])
def test_grid_backward_matches_exact_sum(dtype, grad_dtype, tol, scatter_mode):
    bound = 3
    pls, offsets, emb = _table(bound, dtype, seed=3)
    x = _marched_points(bound, 192, seed=2)[:60000]
    B = len(x)
    rng = np.random.default_rng(0)
    grad = (rng.normal(size=(B, 32)) * 1e-2).astype(dtype)
    g_lbc = np.ascontiguousarray(grad.reshape(B, 16, 2).transpose(1, 0, 2)).astype(np.float32)
    want = oracle.grid_encode_backward(g_lbc, x, offsets, offsets[-1], 2, pls, 16, half_products=(grad_dtype == torch.float16),
                                       level_scales=gpu_level_scales(pls, 16, 16))
    gg = torch.zeros(int(offsets[-1]), 2, device=DEV, dtype=grad_dtype)
    dummy = torch.zeros(1, device=DEV, dtype=t(grad).dtype)
    GB.grid_encode_backward(t(grad), t(x), t(emb), t(offsets), gg, B, 3, 2, 16, np.log2(pls), 16, False, dummy, dummy, 0, 1)
    err = np.abs(n(gg).astype(np.float64) - want).max()
    assert err <= tol * np.abs(want).max(), f"max err {err} vs max |grad| {np.abs(want).max()}"
    # layout 0 ([L,B,C]) gives the same table
    gg0 = torch.zeros_like(gg)
    GB.grid_encode_backward(t(g_lbc.astype(dtype)), t(x), t(emb), t(offsets), gg0, B, 3, 2, 16, np.log2(pls), 16, False, dummy, dummy, 0, 0)
    assert np.abs(n(gg0).astype(np.float64) - want).max() <= tol * np.abs(want).max()
    R = ref_mod("_gridencoder")
    if R is not None:
        rg = torch.zeros(int(offsets[-1]), 2, device=DEV, dtype=t(emb).dtype)
        R.grid_encode_backward(t(g_lbc.astype(dtype)), t(x), t(emb), t(offsets), rg, B, 3, 2, 16, float(np.log2(pls)), 16, False, dummy, dummy, 0)
        ref_err = np.abs(n(rg).astype(np.float64) - want).max()
        ours = np.abs(n(gg).astype(np.float64) - n(rg).astype(np.float64)).max()
        lim = (2e-2 if dtype == np.float16 else 1e-4) * np.abs(want).max()
        assert ours <= 2 * lim, f"ours vs reference {ours}, reference vs exact {ref_err}"


def test_grid_backward_generic_shapes(scatter_mode):
    rng = np.random.default_rng(6)
    for D, C, L, gridtype, log2T in [(2, 2, 4, 0, 19), (3, 1, 5, 0, 12), (3, 4, 6, 0, 14), (3, 8, 3, 1, 10), (2, 4, 8, 1, 8), (3, 2, 32, 0, 10)]:
        pls = 1.3
        offsets = oracle.grid_offsets(D, L, pls, 16, log2T)
        emb = rng.uniform(-1, 1, (offsets[-1], C)).astype(np.float32)
        x = np.sort(rng.random((1999, D)).astype(np.float32), axis=0)          # sorted -> long same-cell runs
        x[7] = 1.5                                                              # one out-of-range sample
        B = len(x)
        grad = rng.normal(size=(L, B, C)).astype(np.float32)
        want = oracle.grid_encode_backward(grad, x, offsets, offsets[-1], C, pls, 16, gridtype=gridtype, level_scales=gpu_level_scales(pls, 16, L))
        dummy = torch.zeros(1, device=DEV)
        for layout in (0, 1):
            g = grad if layout == 0 else np.ascontiguousarray(grad.transpose(1, 0, 2)).reshape(B, L * C)
            gg = torch.zeros(int(offsets[-1]), C, device=DEV)
            GB.grid_encode_backward(t(g), t(x), t(emb), t(offsets), gg, B, D, C, L, np.log2(pls), 16, False, dummy, dummy, gridtype, layout)
            err = np.abs(n(gg) - want).max()
            assert err <= 1e-4 * np.abs(want).max(), (D, C, L, gridtype, layout, err)


def test_grid_input_gradient_matches_oracle_and_reference():
    pls, offsets, emb = _table(1, np.float32, seed=8, scale=1.0)
    rng = np.random.default_rng(4)
    x = rng.random((3000, 3)).astype(np.float32)
    B = len(x)
    grad = rng.normal(size=(16, B, 2)).astype(np.float32)
    out, dy = _fwd(x, emb, offsets, pls, 0, cg=True)
    gi = torch.zeros(B, 3, device=DEV)
    gg = torch.zeros(int(offsets[-1]), 2, device=DEV)
    GB.grid_encode_backward(t(grad), t(x), t(emb), t(offsets), gg, B, 3, 2, 16, np.log2(pls), 16, True, dy, gi, 0, 0)
    want = oracle.grid_input_backward(grad, n(dy).reshape(B, 16, 3, 2))
    assert np.allclose(n(gi), want, rtol=1e-4, atol=1e-3)
    R = ref_mod("_gridencoder")
    if R is not None:
        rgi, rgg = torch.zeros(B, 3, device=DEV), torch.zeros_like(gg)
        R.grid_encode_backward(t(grad), t(x), t(emb), t(offsets), rgg, B, 3, 2, 16, float(np.log2(pls)), 16, True, dy, rgi, 0)
        assert torch.allclose(rgi, gi, rtol=1e-4, atol=1e-3)


def test_grid_module_autograd_and_autocast():
    torch.manual_seed(0)
    enc = gridencoder.GridEncoder(desired_resolution=2048 * 2).to(DEV)
    assert enc.output_dim == 32 and enc.embeddings.shape == (6328848, 2) and enc.offsets.dtype == torch.int32
    with torch.no_grad():
        enc.embeddings.uniform_(-0.5, 0.5)
    x = (torch.rand(5000, 3, device=DEV) * 2 - 1) * 2
    pls, offsets = enc.per_level_scale, n(enc.offsets)
    y = enc(x, bound=2)
    xn = n((x + 2) / 4)
    want, _ = oracle.grid_encode_forward(xn, n(enc.embeddings), offsets, pls, 16, level_scales=gpu_level_scales(pls, 16, 16))
    assert y.dtype == torch.float32 and np.allclose(n(y).reshape(5000, 16, 2).transpose(1, 0, 2), want, atol=1e-6)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    exact = oracle.grid_encode_backward(n(g).reshape(5000, 16, 2).transpose(1, 0, 2), xn, offsets, offsets[-1], 2, pls, 16,
                                        level_scales=gpu_level_scales(pls, 16, 16))
    assert enc.embeddings.grad.dtype == torch.float32
    assert np.abs(n(enc.embeddings.grad) - exact).max() < 1e-4 * np.abs(exact).max()
    enc.embeddings.grad = None
    with torch.autocast("cuda", dtype=torch.float16):
        yh = enc(x, bound=2)
        assert yh.dtype == torch.float16
        (yh.float() * g).sum().backward()
    wanth, _ = oracle.grid_encode_forward(xn, n(enc.embeddings).astype(np.float16), offsets, pls, 16, level_scales=gpu_level_scales(pls, 16, 16))
    assert np.abs(n(yh).astype(np.float32).reshape(5000, 16, 2).transpose(1, 0, 2) - wanth.astype(np.float32)).max() <= 1e-3
    assert enc.embeddings.grad.dtype == torch.float32
    assert np.abs(n(enc.embeddings.grad) - exact).max() < 2e-3 * np.abs(exact).max()       # grad rounded to fp16 on the way in
    # the cached fp16 table follows parameter updates
    with torch.no_grad():
        enc.embeddings.mul_(2.0)
    with torch.autocast("cuda", dtype=torch.float16):
        y2 = enc(x, bound=2)
    assert torch.allclose(y2.float(), 2 * yh.float(), atol=2e-3)


# ------------------------------------------------------------------------------------- SH
@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_matches_oracle_and_reference(degree):
    rng = np.random.default_rng(degree)
    d = rng.normal(size=(4099, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d32 = d.astype(np.float32)
    B, C2 = len(d), degree * degree
    out = torch.empty(B, C2, device=DEV)
    dy = torch.empty(B, 3 * C2, device=DEV)
    SB.sh_encode_forward(t(d32), out, B, 3, degree, True, dy)
    want = oracle.sh_encode(d32, degree) if degree <= 4 else oracle.sh_encode_scipy(d32, degree)
    assert np.abs(n(out) - want).max() < (2e-6 if degree <= 4 else 2e-5)
    R = ref_mod("_shencoder")
    if R is not None:
        # non-unit inputs too: the reference's closed forms are polynomials in x,y,z
        dd = np.concatenate([d32, (rng.normal(size=(512, 3)) * 0.7).astype(np.float32)])
        B2 = len(dd)
        rout, rdy = torch.empty(B2, C2, device=DEV), torch.empty(B2, 3 * C2, device=DEV)
        R.sh_encode_forward(t(dd), rout, B2, 3, degree, True, rdy)
        gout, gdy = torch.empty(B2, C2, device=DEV), torch.empty(B2, 3 * C2, device=DEV)
        SB.sh_encode_forward(t(dd), gout, B2, 3, degree, True, gdy)
        assert torch.allclose(gout, rout, atol=3e-5, rtol=1e-4), float((gout - rout).abs().max())
        assert torch.allclose(gdy, rdy, atol=3e-4, rtol=1e-4), float((gdy - rdy).abs().max())
        # fp16 I/O
        rh, gh = torch.empty(B2, C2, device=DEV, dtype=torch.half), torch.empty(B2, C2, device=DEV, dtype=torch.half)
        dum = torch.empty(1, device=DEV, dtype=torch.half)
        R.sh_encode_forward(t(dd).half(), rh, B2, 3, degree, False, dum)
        SB.sh_encode_forward(t(dd).half(), gh, B2, 3, degree, False, dum)
        exact = torch.empty(B2, C2, device=DEV)
        SB.sh_encode_forward(t(dd).half().float(), exact, B2, 3, degree, False, torch.empty(1, device=DEV))
        assert float((gh.float() - exact).abs().max()) <= float((rh.float() - exact).abs().max()) + 1e-3
        if degree <= 4:     # higher degrees: the reference's fp16 polynomial evaluation itself is off by > 2e-2
            assert torch.allclose(gh.float(), rh.float(), atol=2e-2, rtol=2e-2)


def test_sh_module_forward_backward():
    enc = shencoder.SHEncoder(degree=4)
    d = torch.randn(1000, 3, device=DEV)
    d = (d / d.norm(dim=-1, keepdim=True)).requires_grad_(True)
    y = enc(d)
    assert y.shape == (1000, 16)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    # finite differences of the oracle's closed forms
    eps = 1e-3
    dn, gn = n(d), n(g)
    for k in range(3):
        dp, dm = dn.copy(), dn.copy()
        dp[:, k] += eps
        dm[:, k] -= eps
        fd = ((oracle.sh_encode(dp, 4).astype(np.float64) - oracle.sh_encode(dm, 4)) / (2 * eps) * gn).sum(-1)
        assert np.allclose(n(d.grad)[:, k], fd, atol=5e-3, rtol=1e-2)
    with torch.autocast("cuda", dtype=torch.float16):
        yh = enc(d.detach())
    assert yh.dtype == torch.float16 and torch.allclose(yh.float(), y.detach(), atol=2e-3)


@pytest.mark.parametrize("bound", [1, 2, 3, 0.75])
@pytest.mark.parametrize("autocast", [False, True])
def test_grid_module_input_mapping_inside_the_kernels_is_bit_identical(bound, autocast):
    """GridEncoder.forward hands raw positions to the kernels, which apply (x + bound) / (2 * bound) in ATen's arithmetic (a scalar
    divisor is a multiplication by its fp32 reciprocal): features and table gradient equal, to the bit, those of the explicit
    ATen expression followed by the [0, 1] kernels — including positions outside [-bound, bound] (zero features, no gradient)"""
    torch.manual_seed(4)
    enc = gridencoder.GridEncoder(desired_resolution=2048 * max(1, int(bound))).to(DEV)
    with torch.no_grad():
        enc.embeddings.uniform_(-0.5, 0.5)
    g = torch.Generator(device=DEV).manual_seed(9)
    x = (torch.rand(20011, 3, device=DEV, generator=g) * 2 - 1) * (bound * 1.02)          # 2 % outside
    gout = torch.randn(20011, 32, device=DEV, generator=g)
    res = []
    for fused in (True, False):
        enc.embeddings.grad = None
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            if fused:
                f = enc(x, bound=bound)
            else:
                unit = (x + bound) / (2 * bound)
                f = gridencoder.grid.grid_encode(unit, enc.embeddings, enc.offsets, enc.per_level_scale, enc.base_resolution, False, enc.gridtype_id)
        (f.float() * gout).sum().backward()
        res.append((f.detach().clone(), enc.embeddings.grad.clone()))
    assert torch.equal(res[0][0], res[1][0])
    assert int((res[0][0].float().abs().sum(dim=1) == 0).sum()) > 100                     # the out-of-range rows
    scale = float(res[1][1].abs().max())
    assert float((res[0][1] - res[1][1]).abs().max()) <= 1e-5 * scale                     # same contributions, atomics in another order

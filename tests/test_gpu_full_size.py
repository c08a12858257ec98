"""GPU parity at BASELINE.json's full size (configs[1]: 4096 rays, bound 3 -> ~3.3 M samples).

The oracle needs minutes at this size, so the kernels are checked through properties that do not
depend on the size and tie the full-size launch to the small launches the oracle *does* verify
(tests/test_gpu_encoders.py, tests/test_gpu_ffmlp.py):

  * row independence — a sample's result in the full launch equals its result in a small launch
    of a random subset (bit-exact: same instruction sequence per sample/row);
  * partition of unity — a table of ones encodes to ones (the eight blend weights sum to 1);
  * exact homogeneity — scaling an fp32 table by 4 scales the encoding by exactly 4;
  * checksum of checksums — per level, the sum of the scattered table gradient equals the sum of
    the incoming gradient column (again: weights sum to 1);
  * additivity — the gradient of the whole batch equals the sum of the gradients of its halves.

Tolerances: bit-exact where stated; sums of fp32 atomics in arbitrary order are compared in
float64 with 1e-5 relative (checksums) / 1e-4 of the largest entry (additivity); fp16 MLP outputs
against a float32 torch evaluation of the same network: 2e-3 * max|value| (tests/test_gpu_ffmlp.py).
"""
import numpy as np
import pytest
import torch

from enerf_b200 import raymarching as rm
from enerf_b200.backends import ffmlp_backend as FB
from enerf_b200.backends import gridencoder_backend as GB
from oracle import oracle
from tests.gpu_common import DEV, scene, t

pytestmark = pytest.mark.gpu

BOUND, N_RAYS = 3, 4096


@pytest.fixture(scope="module")
def samples():
    """unit-cube coordinates of every sample the marcher emits for the BASELINE batch, marcher order"""
    sc = scene(N_RAYS, BOUND, seed=11)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, _, _, _ = rm.march_rays_train(t(sc["o"]), t(sc["d"]), float(BOUND), t(sc["bits"]), 3, 128, t(sc["nears"]), t(sc["fars"]), counter,
                                        -1, True, 128, False, 0, 1024)
    m = int(counter[0])
    assert m > 2_000_000, m
    return ((xyzs[:m] + BOUND) / (2 * BOUND)).contiguous()


def _grid():
    pls = oracle.per_level_scale_for(2048 * BOUND, 16, 16)
    offsets = oracle.grid_offsets(3, 16, pls, 16, 19)
    return pls, t(offsets), int(offsets[-1]), offsets


def _encode(x, emb, toff, pls):
    B = x.shape[0]
    out = torch.empty(B, 32, device=DEV, dtype=emb.dtype)
    dummy = torch.empty(1, device=DEV, dtype=emb.dtype)
    GB.grid_encode_forward(x, emb, toff, out, B, 3, 2, 16, np.log2(pls), 16, False, dummy, 0, 1)
    return out


def _scatter(grad, x, emb, toff, n_entries, pls):
    B = x.shape[0]
    gg = torch.zeros(n_entries, 2, device=DEV, dtype=torch.float32)
    dummy = torch.zeros(1, device=DEV, dtype=grad.dtype)
    GB.grid_encode_backward(grad, x, emb, toff, gg, B, 3, 2, 16, np.log2(pls), 16, False, dummy, dummy, 0, 1)
    return gg


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_gather_rows_are_independent_of_the_launch_size(samples, dtype):
    pls, toff, n_entries, _ = _grid()
    g = torch.Generator(device=DEV).manual_seed(1)
    emb = (torch.rand(n_entries, 2, device=DEV, generator=g) * 2 - 1).to(dtype)
    full = _encode(samples, emb, toff, pls)
    pick = torch.randperm(samples.shape[0], device=DEV, generator=g)[:8191]          # ragged on purpose
    small = _encode(samples[pick].contiguous(), emb, toff, pls)
    assert torch.equal(full[pick], small)
    assert torch.isfinite(full.float()).all()


def test_gather_partition_of_unity_and_homogeneity(samples):
    pls, toff, n_entries, _ = _grid()
    ones = torch.ones(n_entries, 2, device=DEV)
    enc = _encode(samples, ones, toff, pls)
    assert float((enc - 1).abs().max()) < 1e-5
    g = torch.Generator(device=DEV).manual_seed(2)
    emb = torch.rand(n_entries, 2, device=DEV, generator=g) * 2 - 1
    assert torch.equal(_encode(samples, emb * 4, toff, pls), _encode(samples, emb, toff, pls) * 4)


def test_scatter_checksums_and_additivity(samples):
    pls, toff, n_entries, offsets = _grid()
    B = samples.shape[0]
    g = torch.Generator(device=DEV).manual_seed(3)
    emb = torch.zeros(n_entries, 2, device=DEV, dtype=torch.float16)
    grad = (torch.randn(B, 32, device=DEV, generator=g).abs() * 1e-2).half()
    gg = _scatter(grad, samples, emb, toff, n_entries, pls)
    col = grad.double().sum(0).view(16, 2)
    for lv in range(16):
        got = gg[offsets[lv]:offsets[lv + 1]].double().sum(0)
        assert torch.allclose(got, col[lv], rtol=1e-5, atol=0), (lv, got.tolist(), col[lv].tolist())
    h = (B // 2) // 128 * 128 + 37                                                   # ragged split
    parts = _scatter(grad[:h].contiguous(), samples[:h].contiguous(), emb, toff, n_entries, pls) + \
        _scatter(grad[h:].contiguous(), samples[h:].contiguous(), emb, toff, n_entries, pls)
    assert float((gg - parts).abs().max()) <= 1e-4 * float(gg.abs().max())


def _net(I, W, nl, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    nw = W * (I + W * (nl - 1) + 16)
    return ((torch.rand(nw, device=DEV, generator=g) * 2 - 1) * (3 / W) ** 0.5).half()


def _torch_mlp(x, w, I, W, nl):
    """float32 evaluation with fp16 storage of the hidden activations (what forward_buffer holds)"""
    h, pos = x.float(), 0
    for layer in range(nl + 1):
        rows = 16 if layer == nl else W
        cols = I if layer == 0 else W
        m = w[pos:pos + rows * cols].view(rows, cols).float()
        pos += rows * cols
        h = h @ m.t()
        if layer < nl:
            h = torch.relu(h).half().float()
    return h


@pytest.mark.parametrize("nl", [2, 3])
def test_mlp_rows_are_independent_of_the_launch_size(samples, nl):
    I, W = 32, 64
    B = samples.shape[0] // 128 * 128
    g = torch.Generator(device=DEV).manual_seed(4 + nl)
    x = (torch.randn(B, I, device=DEV, generator=g) * 0.5).half()
    w = _net(I, W, nl, seed=nl)
    out = torch.empty(B, 16, device=DEV, dtype=torch.half)
    FB.ffmlp_inference(x, w, B, I, 16, W, nl, 0, 6, None, out)
    want = _torch_mlp(x, w, I, W, nl)
    assert float((out.float() - want).abs().max()) <= 2e-3 * float(want.abs().max()) + 1e-3
    lo = 128 * 1000
    small = torch.empty(128 * 33, 16, device=DEV, dtype=torch.half)
    FB.ffmlp_inference(x[lo:lo + 128 * 33].contiguous(), w, 128 * 33, I, 16, W, nl, 0, 6, None, small)
    assert torch.equal(out[lo:lo + 128 * 33], small)

    # backward without a forward_buffer (the training path): input gradients are per row, weight gradients add up
    grad = (torch.randn(B, 16, device=DEV, generator=g) * 0.1).half()

    def bwd(xs, gs):
        b = xs.shape[0]
        gin = torch.empty(b, I, device=DEV, dtype=torch.half)
        gw = torch.zeros(w.numel(), device=DEV, dtype=torch.float32)
        FB.ffmlp_backward(gs, xs, w, None, b, I, 16, W, nl, 0, 6, True, None, gin, gw)
        return gin, gw

    gin, gw = bwd(x, grad)
    h = (B // 2) // 128 * 128
    gin_a, gw_a = bwd(x[:h].contiguous(), grad[:h].contiguous())
    gin_b, gw_b = bwd(x[h:].contiguous(), grad[h:].contiguous())
    assert torch.equal(gin[:h], gin_a) and torch.equal(gin[h:], gin_b)
    assert float((gw - (gw_a + gw_b)).abs().max()) <= 1e-4 * float(gw.abs().max())
    # against autograd through the float32 evaluation: ReLU units within rounding of zero may flip -> relative L2
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    h32, pos = xr, 0
    for layer in range(nl + 1):
        rows = 16 if layer == nl else W
        cols = I if layer == 0 else W
        m = wr[pos:pos + rows * cols].view(rows, cols)
        pos += rows * cols
        h32 = h32 @ m.t()
        if layer < nl:
            h32 = torch.relu(h32)
    h32.backward(grad.float())
    rel_w = float((gw - wr.grad).norm() / wr.grad.norm())
    rel_x = float((gin.float() - xr.grad).norm() / xr.grad.norm())
    assert rel_w < 2e-2 and rel_x < 2e-2, (rel_w, rel_x)

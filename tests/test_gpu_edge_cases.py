"""Edge cases of the hot path on the GPU: empty inputs, rays that miss the volume, ragged sizes around the tile
granularities (32-sample warps, 128-sample MLP tiles), out-of-range samples.  The reference's wrappers accept all of these
(they only launch ceil-div grids), so the drop-in must too."""
import numpy as np
import pytest
import torch

from enerf_b200 import gridencoder, raymarching as rm, shencoder, synthetic
from enerf_b200.ffmlp import FFMLP
from enerf_b200.nerf.network_ff import NeRFNetwork as FFNet
from oracle import oracle
from tests.gpu_common import DEV, n, t

pytestmark = pytest.mark.gpu


def test_empty_inputs_everywhere():
    z3 = torch.zeros(0, 3, device=DEV)
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1], device=DEV)
    nears, fars = rm.near_far_from_aabb(z3, z3, aabb, 0.2)
    assert nears.shape == (0,) and fars.shape == (0,)
    enc = gridencoder.GridEncoder().to(DEV)
    out = enc(z3, bound=1)
    assert out.shape == (0, 32)
    sh = shencoder.SHEncoder().to(DEV)
    assert sh(z3).shape == (0, 16)
    ws, dp, im = rm.composite_rays_train(torch.zeros(0, device=DEV), torch.zeros(0, 3, device=DEV), torch.zeros(0, 2, device=DEV),
                                         torch.zeros(0, 3, dtype=torch.int32, device=DEV))
    assert ws.shape == (0,) and im.shape == (0, 3)
    mlp = FFMLP(32, 16, 64, 2).to(DEV)
    with torch.autocast("cuda", dtype=torch.float16):
        y = mlp(torch.zeros(0, 32, device=DEV))
    assert y.shape == (0, 16)


def test_rays_that_miss_the_volume_render_background():
    bound = 2
    torch.manual_seed(0)
    model = FFNet(bound=bound, cuda_ray=True, out_dim_color=1).to(DEV)
    model.density_bitfield.copy_(t(synthetic.packbits_np(synthetic.ball_density_grid(bound, model.cascade))))
    # rays starting far outside and pointing away from the box, mixed with rays that hit it
    o_hit, d_hit = synthetic.random_rays(64, bound, seed=1)
    o_miss = np.tile(np.array([[10.0, 10.0, 10.0]], np.float32), (64, 1))
    d_miss = np.tile(np.array([[0.577, 0.577, 0.577]], np.float32), (64, 1))
    o = np.concatenate([o_hit, o_miss])
    d = np.concatenate([d_hit, d_miss])
    for mode in ("train", "eval"):
        getattr(model, mode)()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out = model.render(t(o)[None], t(d)[None], bg_color=1, perturb=False, out_dim_color=1)
        img = n(out["image"]).reshape(-1)
        assert np.isfinite(img[:64]).all()
        assert np.allclose(img[64:], 1.0), f"{mode}: a ray that misses the volume must show the background"


@pytest.mark.parametrize("B", [1, 31, 33, 127, 129, 4097])
def test_ragged_batch_sizes(B):
    """grid encoder / SH / FFMLP on batch sizes around the warp and tile granularities, vs the oracle."""
    rng = np.random.default_rng(B)
    x = rng.random((B, 3)).astype(np.float32)
    pls = oracle.per_level_scale_for(2048, 16, 16)
    offsets = oracle.grid_offsets(3, 16, pls, 16, 19)
    emb = rng.uniform(-1, 1, (offsets[-1], 2)).astype(np.float32)
    enc = gridencoder.GridEncoder(desired_resolution=2048).to(DEV)
    with torch.no_grad():
        enc.embeddings.copy_(t(emb))
    got = n(enc(t(x) * 2 - 1, bound=1))              # module maps [-1,1] -> [0,1]
    lv = torch.arange(16, device=DEV, dtype=torch.float32)
    scales = n(torch.exp2(lv * float(np.float32(np.log2(pls)))) * 16.0 - 1.0).astype(np.float32)
    x01 = n((t(x) * 2 - 1 + 1) / 2)
    want, _ = oracle.grid_encode_forward(x01, emb, offsets, pls, 16, level_scales=scales)
    assert np.allclose(got.reshape(B, 16, 2).transpose(1, 0, 2), want, atol=1e-6)
    # FFMLP pads to a multiple of 128 internally (ffmlp.py:157-159) and slices back
    torch.manual_seed(1)
    mlp = FFMLP(32, 16, 64, 2).to(DEV)
    xin = (torch.randn(B, 32, device=DEV) * 0.5)
    with torch.autocast("cuda", dtype=torch.float16):
        y = mlp(xin)
    w = n(mlp.weights).astype(np.float16)
    ywant, _ = oracle.ffmlp_forward(n(xin.half()), w, 32, 64, 2)
    assert y.shape == (B, 16)
    assert np.abs(n(y).astype(np.float64) - ywant).max() <= 2e-3 * np.abs(ywant).max() + 1e-4


def test_out_of_range_samples_encode_to_zero_and_get_no_gradient():
    enc = gridencoder.GridEncoder().to(DEV)
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    x = torch.tensor([[0.0, 0.0, 0.0], [1.5, 0.0, 0.0], [0.2, -1.2, 0.3], [0.9999, 0.9999, 0.9999]], device=DEV)
    out = enc(x, bound=1)
    assert torch.all(out[1] == 0) and torch.all(out[2] == 0) and out[0].abs().sum() > 0 and out[3].abs().sum() > 0
    w = torch.zeros(4, 1, device=DEV)
    w[1] = w[2] = 1.0                                # only the out-of-range samples carry gradient
    (out * w).sum().backward()
    assert float(enc.embeddings.grad.abs().sum()) == 0.0

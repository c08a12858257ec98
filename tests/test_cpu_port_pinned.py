"""Pins `oracle/cpu_reference.py` — the CPU port that `bench.py --impl reference`, `cpu_baseline` and the PSNR-parity run use
on the GPU box, where /root/reference does not exist — to the reference's OWN Python classes (nerf/network.py NeRFNetwork +
nerf/renderer.py NeRFRenderer.run, imported unchanged through oracle/ref_python.py) on identical parameters and rays:
image, depth and every parameter gradient to 1e-6.  Build container only (skipped where the reference tree is absent)."""
import numpy as np
import pytest
import torch

from enerf_b200 import synthetic
from oracle import cpu_reference, ref_python

pytestmark = pytest.mark.skipif(not ref_python.available(), reason="reference tree only exists in the build container")


def _pair(bound, n_ch, seed):
    ns = ref_python.load()
    torch.manual_seed(seed)
    ref = ns.make_network(encoding="hashgrid", bound=bound, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1,
                          out_dim_color=n_ch)
    with torch.no_grad():
        ref.encoder.embeddings.uniform_(-0.3, 0.3)
    port = cpu_reference.NeRFNetworkCPU(bound=bound, out_dim_color=n_ch)
    with torch.no_grad():
        port.encoder.embeddings.copy_(ref.encoder.embeddings)
        for a, b in zip(port.sigma_net, ref.sigma_net):
            a.weight.copy_(b.weight)
        for a, b in zip(port.color_net, ref.color_net):
            a.weight.copy_(b.weight)
    return ref, port


def _grads(model):
    return [model.encoder.embeddings.grad] + [l.weight.grad for l in model.sigma_net] + [l.weight.grad for l in model.color_net]


@pytest.mark.parametrize("bound,n_ch,upsample", [(1, 3, 0), (3, 1, 0), (2, 1, 24)])
def test_cpu_port_matches_the_reference_classes(bound, n_ch, upsample):
    ref, port = _pair(bound, n_ch, seed=bound)
    if upsample:            # deterministic inverse-CDF samples (det = not training), so both sides draw the same new z values
        ref.eval()
        port.eval()
    o, d = synthetic.random_rays(96, bound, seed=11)
    o, d = torch.from_numpy(o), torch.from_numpy(d)
    target = torch.rand(96, n_ch, generator=torch.Generator().manual_seed(1))
    out_r = ref.render(o[None], d[None], staged=False, num_steps=48, upsample_steps=upsample, bg_color=1, perturb=False, out_dim_color=n_ch)
    out_p = port.render(o, d, num_steps=48, bg_color=1, perturb=False, upsample_steps=upsample)
    img_r, dep_r = out_r["image"][0], out_r["depth"][0]
    assert float((img_r - out_p["image"]).abs().max()) <= 1e-6
    assert float((dep_r - out_p["depth"]).abs().max()) <= 1e-6
    ((img_r - target) ** 2).mean().backward()
    ((out_p["image"] - target) ** 2).mean().backward()
    for g_r, g_p in zip(_grads(ref), _grads(port)):
        assert g_r is not None and g_p is not None
        assert float((g_r - g_p).abs().max()) <= 1e-6 * max(1.0, float(g_r.abs().max()))
    assert float(_grads(ref)[0].abs().sum()) > 0


def test_cpu_port_perturbed_sampling_matches_with_the_same_generator_state():
    """perturb=True: both sides draw `torch.rand(z_vals.shape)` once (renderer.py:178-179)"""
    ref, port = _pair(1, 1, seed=5)
    o, d = synthetic.random_rays(64, 1, seed=12)
    o, d = torch.from_numpy(o), torch.from_numpy(d)
    torch.manual_seed(77)
    out_r = ref.render(o[None], d[None], staged=False, num_steps=32, upsample_steps=0, bg_color=1, perturb=True, out_dim_color=1)
    torch.manual_seed(77)
    out_p = port.render(o, d, num_steps=32, bg_color=1, perturb=True)
    assert float((out_r["image"][0] - out_p["image"]).abs().max()) <= 1e-6
    assert np.isfinite(out_p["depth"].detach().numpy()).all()

"""Host-side mirror checks that need no GPU."""
from enerf_b200.nerf.network import NeRFNetwork


def test_mirror_state_dict_matches_the_reference_layout():
    """checkpoints stay interchangeable: parameter / buffer names and shapes of both mirrors"""
    from enerf_b200.nerf.network_ff import NeRFNetwork as FFNet
    sd = NeRFNetwork(bound=2, cuda_ray=True, out_dim_color=1).state_dict()
    assert sd["encoder.embeddings"].shape[1] == 2 and sd["encoder.offsets"].shape == (17,)
    assert sd["sigma_net.0.weight"].shape == (64, 32) and sd["sigma_net.1.weight"].shape == (16, 64)
    assert sd["color_net.0.weight"].shape == (64, 31) and sd["color_net.2.weight"].shape == (1, 64)
    for k, shape in (("aabb_train", (6,)), ("aabb_infer", (6,)), ("density_grid", (2, 128 ** 3)), ("density_bitfield", (2 * 128 ** 3 // 8,)), ("step_counter", (16, 2))):
        assert tuple(sd[k].shape) == shape, k
    assert not any(k.startswith("_") or "mean_density" in k for k in sd)
    sd = FFNet(bound=2, cuda_ray=True, out_dim_color=1).state_dict()
    assert sd["sigma_net.weights"].shape == (64 * (32 + 64 + 16),) and sd["color_net.weights"].shape == (64 * (32 + 128 + 16),)

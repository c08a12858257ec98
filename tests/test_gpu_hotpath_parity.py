"""Training-level parity of the benchmarked hot path (cuda_ray + ff + fp16 autocast + GradScaler + FusedAdam, tcgen05 fused field)
against the reference's OWN CUDA kernels chained as nerf/network_ff.py + NeRFRenderer.run_cuda (oracle/ref_chain.py, built from the
unmodified sources by oracle/build_ref.py) — the north star's "rendered PSNR within 0.1 dB of the reference".

What can be asserted in a test that runs in seconds:
  * one training step on identical samples: image and depth agree to 1e-4, MLP weight gradients to 2e-2 relative L2, the hash-table
    gradient to 0.15 relative L2 (the reference accumulates it with fp16 atomics, gridencoder.cu:296-302; this repo in fp32);
  * a 600-step training of the tiny scene from identical parameters: both stacks learn it (> 36 dB) and this repo's does not end more
    than 2.5 dB below the reference kernels' (the run-to-run spread of EITHER stack at 600 steps is about +-1.5 dB).
Training on this scene is chaotic and, because of atomics, not reproducible run to run: the same stack lands +-1 dB apart between
repetitions (profiles/r2_13_hotpath_parity.json: 5 seeds x 1500 steps, ours 50.55 +- 0.97 dB, reference kernels 50.85 +- 1.10 dB, mean
difference -0.30 +- 0.34 dB (standard error) — statistically indistinguishable), so a single run cannot resolve 0.1 dB."""
import pytest

from tests.gpu_common import DEV  # noqa: F401

pytestmark = pytest.mark.gpu


def _need_ref():
    from oracle import ref_chain
    if not ref_chain.available():
        pytest.skip("oracle/_ref (the reference's own CUDA build) is not present")


def test_composed_gradient_matches_reference_kernels():
    _need_ref()
    from tests import hotpath_parity
    r = hotpath_parity.gradient_check()
    assert r["samples_ours"] == r["samples_reference"] > 100000, r          # the two marchers emitted the same samples
    assert r["image_max_abs_diff"] <= 1e-4 and r["depth_max_abs_diff"] <= 1e-4, r
    assert r["grad_sigma_net_rel_l2"] <= 2e-2 and r["grad_color_net_rel_l2"] <= 2e-2, r
    assert r["grad_embeddings_norm"] > 0 and r["grad_embeddings_rel_l2"] <= 0.15, r


def test_training_psnr_parity_vs_reference_kernels():
    _need_ref()
    from tests import hotpath_parity
    r = hotpath_parity.run(steps=600, n_rays=1024)
    assert r["psnr_ours_db"] > 36.0 and r["psnr_reference_kernels_db"] > 36.0, r
    assert r["psnr_ours_db"] >= r["psnr_reference_kernels_db"] - 2.5, r
    assert r["psnr_between_db"] > 34.0, r


def test_rendered_psnr_of_one_checkpoint_matches_reference_inference_kernels():
    """north star: "rendered PSNR within 0.1 dB of the reference" — one trained checkpoint through both inference paths"""
    _need_ref()
    from tests import hotpath_parity
    r = hotpath_parity.render_parity(steps=400)
    assert r["psnr_ours_db"] > 20.0, r                                       # a trained checkpoint, not noise
    assert r["abs_diff_db"] <= 0.1, r
    assert r["psnr_between_db"] >= 45.0, r

"""Device-side mask generators: same law as the reference's generators.  The fixture holds
inclusion frequencies of 6000 masks drawn by the REAL reference and the complete set of its equispaced patterns."""
import math

import pytest
import torch

from conftest import load_golden


def _check(dev):
    from deepinv_b200.physics.generator import EquispacedMaskGenerator, GaussianMaskGenerator, RandomMaskGenerator

    g = load_golden("maskgen_stats")
    W, N_ref, N = 64, int(g["n_rows"]), 20000
    for tag, cls, acc in [("random4", RandomMaskGenerator, 4), ("gauss4", GaussianMaskGenerator, 4), ("gauss8", GaussianMaskGenerator, 8)]:
        gen = cls((2, 8, W), acceleration=acc, rng=torch.Generator(device=dev).manual_seed(1), device=dev)
        m = gen.step(N)["mask"]
        assert m.shape == (N, 2, 8, W) and m.device.type == torch.device(dev).type
        assert bool(((m == 0) | (m == 1)).all()) and torch.equal(m[:, 0], m[:, 1]) and torch.equal(m[:, :, :1].expand_as(m), m)
        cols = m[:, 0, 0].cpu()
        assert torch.equal(cols.sum(-1).unique(), g[f"count_{tag}"])           # exactly n_center + n_lines columns
        p_ref, p = g[f"freq_{tag}"].double(), cols.double().mean(0)
        assert torch.equal(p == 1, p_ref == 1)                                  # the centre band
        sigma = torch.sqrt((p_ref * (1 - p_ref)).clamp_min(1e-4) * (1 / N + 1 / N_ref))
        assert float(((p - p_ref).abs() / sigma).max()) < 5.0                   # same inclusion law (5 sigma over 64 columns)
    gen = EquispacedMaskGenerator((2, 6, 8, W), acceleration=4, rng=torch.Generator(device=dev).manual_seed(2), device=dev)
    m = gen.step(256)["mask"]
    assert m.shape == (256, 2, 6, 8, W)
    pats = torch.unique(m[:, 0, :, 0].cpu(), dim=0)
    assert torch.equal(pats, g["equi_patterns"])                                # every offset, sheared across time, bit-exact
    one = RandomMaskGenerator((32, 32), acceleration=4, device=dev, rng=torch.Generator(device=dev).manual_seed(0)).step(0)["mask"]
    assert one.shape == (1, 32, 32)
    with pytest.raises(ValueError):
        RandomMaskGenerator((2, 32, 32), acceleration=4, center_fraction=0.5, device=dev)


def test_motion_blur_generator_matches_reference_draws():
    """same generator seed on CPU -> the reference's PSFs (tests/golden/motionblur_psf.npz, produced by the real reference)"""
    from conftest import rel_err

    from deepinv_b200.physics.generator import MotionBlurGenerator

    g = load_golden("motionblur_psf")
    mb = MotionBlurGenerator((31, 31), rng=torch.Generator().manual_seed(0))
    f = mb.step(3)["filter"]
    assert f.shape == (3, 1, 31, 31) and rel_err(f, g["filt"]) < 1e-6
    f2 = mb.step(2, sigma=0.4, l=0.5, seed=7)["filter"]
    assert rel_err(f2, g["filt_l"]) < 1e-6
    assert torch.allclose(f.sum(dim=(-2, -1)), torch.ones(3, 1), atol=1e-5)


def test_mask_generators_cpu():
    _check("cpu")

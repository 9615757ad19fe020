"""The property checks the reference's own physics tests make (deepinv/tests/test_physics.py: adjointness :717-745, spectral norm
:882-926, pseudo-inverse :946-968, MRI zero pattern :996-1091, Blur == BlurFFT :1338-1381, tomography option grid :1416-1476,
state_dict / clone / composition / differentiability :1733-2125), re-run on the drop-in classes at the reference's test sizes
(find_operator, :121-575) over the emulated kernels."""
import itertools

import pytest
import torch


@pytest.fixture(autouse=True)
def emul_backend(monkeypatch):
    from emul_util import emul_lib

    from deepinv_b200 import ops

    lib = emul_lib()

    def check(rc):
        assert rc == 0, lib.dinvk_last_error()

    monkeypatch.setattr(ops, "_require_cuda", lambda *ts: torch.device("cpu"))
    monkeypatch.setattr(ops, "_stream", lambda dev: None)
    monkeypatch.setattr(ops, "get_lib", lambda: lib)
    monkeypatch.setattr(ops, "check", check)
    ops._ws_cache.clear()
    yield
    ops._ws_cache.clear()


def _find_operator(name):
    """the reference's test operators (test_physics.py:121-575), same sizes and parameters"""
    import deepinv_b200 as dinv
    from deepinv_b200.physics import functional as dF

    P = dinv.physics
    g = torch.Generator().manual_seed(0)
    if name == "MRI":
        return P.MRI(mask=(torch.rand(17, 11, generator=g) > 0.5).float(), img_size=(2, 17, 11)), (2, 17, 11)
    if name == "MultiCoilMRI":
        maps = torch.randn(1, 7, 17, 11, generator=g, dtype=torch.complex64)
        maps = maps / maps.abs().pow(2).sum(1, keepdim=True).sqrt()
        return P.MultiCoilMRI(mask=(torch.rand(17, 11, generator=g) > 0.3).float(), coil_maps=maps, img_size=(2, 17, 11)), (2, 17, 11)
    if name == "DynamicMRI":
        return P.DynamicMRI(mask=(torch.rand(4, 17, 11, generator=g) > 0.5).float(), img_size=(2, 4, 17, 11)), (2, 4, 17, 11)
    if name == "2DParallelBeamCT":
        return P.Tomography(angles=16, img_width=16, normalize=False), (1, 16, 16)
    if name == "fan_beam_CT":
        return P.Tomography(angles=16, img_width=16, normalize=False, fan_beam=True, fan_parameters={"n_detector_pixels": 24}), (1, 16, 16)
    if name.startswith("blur_"):
        filt = dF.gaussian_blur(sigma=(0.25, 0.1), angle=45.0)
        return P.Blur(filter=filt, padding=name[5:]), (3, 17, 19)
    if name == "blurFFT":
        return P.BlurFFT(img_size=(3, 17, 19), filter=dF.bicubic_filter()), (3, 17, 19)
    if name == "down_bicubic":
        return P.Downsampling(img_size=(3, 16, 20), filter="bicubic", factor=2, padding="circular"), (3, 16, 20)
    raise ValueError(name)


OPERATORS = ["MRI", "MultiCoilMRI", "DynamicMRI", "2DParallelBeamCT", "fan_beam_CT", "blur_valid", "blur_circular", "blur_reflect",
             "blur_replicate", "blur_constant", "blurFFT", "down_bicubic"]


@pytest.mark.parametrize("name", OPERATORS)
def test_adjointness_and_autograd_adjoint(name):
    phys, shape = _find_operator(name)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, *shape, generator=g)
    assert abs(complex(phys.adjointness_test(x)).real) < 1e-3            # test_physics.py:717-745
    y = phys.A(x)
    v = torch.randn(y.shape, generator=g)
    xr = x.clone().requires_grad_(True)
    (grad,) = torch.autograd.grad((phys.A(xr) * v).sum(), xr)            # the autograd adjoint is A^T (kernels, not a tape)
    ref = phys.A_adjoint(v)
    assert float((grad - ref).norm() / ref.norm()) < 1e-5


@pytest.mark.parametrize("name,expected", [("MRI", 1.0), ("DynamicMRI", 1.0), ("MultiCoilMRI", 1.0), ("blurFFT", 1.0), ("blur_circular", 1.0)])
def test_spectral_norm_by_power_method(name, expected):
    phys, shape = _find_operator(name)
    x0 = torch.randn(1, *shape, generator=torch.Generator().manual_seed(2))
    norm = float(phys.compute_norm(x0, max_iter=60, tol=1e-4, verbose=False))
    assert abs(norm - expected) < 0.05, norm                              # test_physics.py:882-926 (1e-2 .. 5e-2 there)


@pytest.mark.parametrize("name", ["MRI", "MultiCoilMRI", "blurFFT"])
def test_pseudo_inverse_reproduces_the_measurements(name):
    phys, shape = _find_operator(name)
    x = torch.randn(1, *shape, generator=torch.Generator().manual_seed(3))
    r = phys.A(x)
    y = phys.A(phys.A_dagger(r))                                          # A A^+ r = r on the range of A
    assert float((y - r).norm() / r.norm()) < 0.05                        # test_physics.py:946-968 (5 %)


def test_projections_and_blur_equals_blurfft():
    phys, shape = _find_operator("blurFFT")
    x = torch.randn(2, *shape, generator=torch.Generator().manual_seed(4))
    assert float((phys.V(phys.V_adjoint(x)) - x).norm() / x.norm()) < 1e-5   # V V^T = I on images (test_physics.py:971-988)
    k = phys.U_adjoint(x)
    assert float((phys.U_adjoint(phys.U(k)) - k).norm() / k.norm()) < 1e-5
    import deepinv_b200 as dinv
    from deepinv_b200.physics import functional as dF

    for (H, W), hw in itertools.product([(17, 19), (16, 18)], [(3, 3), (4, 5), (5, 4)]):   # odd / even image and filter sizes
        filt = torch.rand(1, 1, *hw, generator=torch.Generator().manual_seed(H + hw[0]))
        filt = filt / filt.sum()
        xi = torch.randn(2, 2, H, W, generator=torch.Generator().manual_seed(5))
        a = dinv.physics.Blur(filter=filt, padding="circular").A(xi)
        b = dinv.physics.BlurFFT(img_size=(2, H, W), filter=filt).A(xi)
        assert torch.allclose(a, b, atol=1e-5), (H, W, hw)                # test_blur, test_physics.py:1338-1381


@pytest.mark.parametrize("circle,via_backprop,boundary,normalize", list(itertools.product([False, True], [False, True], [False, True], [False, True])))
def test_tomography_option_grid(circle, via_backprop, boundary, normalize):
    """test_physics.py:1416-1476: every option combination keeps A / A^T consistent and, when normalised, of unit norm"""
    import warnings

    import deepinv_b200 as dinv

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        phys = dinv.physics.Tomography(angles=16, img_width=16, circle=circle, adjoint_via_backprop=via_backprop,
                                       fbp_interpolate_boundary=boundary, normalize=normalize)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 1, 16, 16, generator=g)
    if circle:  # the operator acts on the inscribed disc
        ax = 2 * torch.arange(16.0) / 15 - 1
        x = x * ((ax[None] ** 2 + ax[:, None] ** 2) <= 1)
    y = phys.A(x)
    v = torch.randn(y.shape, generator=g)
    lhs, rhs = float((y * v).sum()), float((x * phys.A_adjoint(v)).sum())
    if via_backprop:  # exact transpose; the IRadon "adjoint" (adjoint_via_backprop=False) is only approximate in the reference too
        assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))
    if normalize:
        n = float(phys.compute_norm(torch.randn(1, 1, 16, 16, generator=g), max_iter=80, tol=1e-5, verbose=False))
        assert abs(n - 1.0) < (1e-2 if via_backprop else 0.3), n
    rec = phys.A_dagger(y, fbp=True)
    assert rec.shape == x.shape and torch.isfinite(rec).all()


def test_state_dict_clone_and_mask_update_semantics():
    import deepinv_b200 as dinv

    phys, shape = _find_operator("MRI")
    x = torch.randn(2, *shape, generator=torch.Generator().manual_seed(7))
    sd = phys.state_dict()
    assert "mask" in sd                                                   # buffers stay registered buffers (test_physics.py:1952-2005)
    other = dinv.physics.MRI(img_size=shape)
    other.load_state_dict(sd)
    assert torch.equal(other.A(x), phys.A(x))
    twin = phys.clone()
    twin.update(mask=torch.ones(17, 11))
    assert not torch.equal(twin.mask, phys.mask) and float(phys.mask.mean()) < 1.0   # clone is independent
    y = phys.A(x)
    assert torch.all((y == 0) == (phys.mask == 0))                        # exact zero pattern (test_physics.py:996-1091)
    new_mask = (torch.rand(17, 11, generator=torch.Generator().manual_seed(8)) > 0.7).float()
    y2 = phys(x, mask=new_mask)                                           # forward stores the mask
    assert torch.all((y2 == 0) == (phys.mask == 0)) and torch.equal(phys.mask[0, 0], new_mask)
    assert phys.A_adjoint(y2, mag=True).shape == (2, 1, 17, 11)


def test_composition_is_linear_and_adjoint_consistent():
    blur, shape = _find_operator("blur_circular")
    import deepinv_b200 as dinv

    down = dinv.physics.Downsampling(img_size=shape, filter=None, factor=1, padding="circular")
    comp = down * blur
    x = torch.randn(1, *shape, generator=torch.Generator().manual_seed(9))
    assert torch.allclose(comp.A(x), blur.A(x), atol=1e-6)
    assert abs(complex(comp.adjointness_test(x)).real) < 1e-3

"""Edge cases the reference's own tests exercise on the package's public API over the emulated kernels: empty
batches, degenerate / odd / prime sizes (which take the mixed-radix and O(N^2) transform paths), non-contiguous and float64
inputs, and the error types of shape / batch mismatches."""
import pytest
import torch

from conftest import rel_err


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


def _operators():
    import deepinv_b200 as dinv

    P = dinv.physics
    filt = torch.ones(1, 1, 3, 3) / 9
    maps = torch.ones(1, 2, 8, 8, dtype=torch.complex64) / 2 ** 0.5
    return [
        ("MRI", P.MRI(img_size=(2, 8, 8)), (2, 8, 8)),
        ("MultiCoilMRI", P.MultiCoilMRI(coil_maps=maps, img_size=(2, 8, 8)), (2, 8, 8)),
        ("DynamicMRI", P.DynamicMRI(mask=torch.ones(3, 8, 8), img_size=(2, 3, 8, 8)), (2, 3, 8, 8)),
        ("Blur", P.Blur(filter=filt, padding="reflect"), (1, 8, 8)),
        ("BlurFFT", P.BlurFFT(img_size=(1, 8, 8), filter=filt), (1, 8, 8)),
        ("Downsampling", P.Downsampling(img_size=(1, 8, 8), filter="bilinear", factor=2), (1, 8, 8)),
        ("Tomography", P.Tomography(angles=4, img_width=8, normalize=False), (1, 8, 8)),
        ("Tomography fan", P.Tomography(angles=4, img_width=8, normalize=False, fan_beam=True,
                                        fan_parameters={"n_detector_pixels": 9}), (1, 8, 8)),
    ]


def test_empty_batches_pass_through_every_operator():
    for name, phys, shape in _operators():
        x = torch.zeros(0, *shape)
        y = phys.A(x)
        assert y.shape[0] == 0, name
        xt = phys.A_adjoint(y)
        assert tuple(xt.shape) == (0, *shape), name
    import deepinv_b200 as dinv

    den = dinv.models.DnCNN(in_channels=1, out_channels=1, depth=3, nf=4, pretrained=None)
    with torch.no_grad():
        assert den(torch.zeros(0, 1, 8, 8), 0.1).shape == (0, 1, 8, 8)


@pytest.mark.parametrize("H,W", [(1, 1), (1, 7), (7, 1), (2, 3), (13, 17), (31, 64), (100, 36)])
def test_mri_degenerate_and_prime_sizes(H, W):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    gen = torch.Generator().manual_seed(H * 131 + W)
    x = torch.randn(2, 2, H, W, generator=gen)
    m = (torch.rand(2, 2, H, W, generator=gen) > 0.4).float()
    p = dinv.physics.MRI(mask=m, img_size=(2, H, W))
    err = lambda a, b: float((a - b).abs().max()) if H * W == 1 else rel_err(a, b)
    y = R.mri_A(x, m)
    assert err(p.A(x), y) < 1e-6 and torch.equal(p.A(x) == 0, y == 0)
    assert err(p.A_adjoint(y), R.mri_At(y, m)) < 1e-6
    assert err(p.prox_l2(x, y, 0.9), R.mri_prox_l2(x, y, m, 0.9)) < 1e-5


def test_non_contiguous_float64_and_integer_like_inputs():
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    gen = torch.Generator().manual_seed(5)
    m = (torch.rand(1, 2, 24, 16, generator=gen) > 0.5).float()
    p = dinv.physics.MRI(mask=m, img_size=(2, 24, 16))
    xt = torch.randn(2, 2, 16, 24, generator=gen).transpose(-1, -2)  # a transposed view, like the sinograms the reference hands out
    assert not xt.is_contiguous() and rel_err(p.A(xt), R.mri_A(xt.contiguous(), m)) < 1e-6
    xs = torch.randn(4, 2, 24, 16, generator=gen)[::2]                 # batch-strided view
    assert rel_err(p.A(xs), R.mri_A(xs.contiguous(), m)) < 1e-6
    xd = torch.randn(2, 2, 24, 16, generator=gen, dtype=torch.float64)  # computed in fp32 (the kernels' arithmetic type)
    y = p.A(xd)
    assert y.dtype == torch.float32 and rel_err(y.double(), R.mri_A(xd, m.double())) < 1e-6
    blur = dinv.physics.Blur(filter=torch.ones(1, 1, 3, 3) / 9, padding="circular")
    img = torch.randn(1, 3, 10, 12, generator=gen).permute(0, 1, 3, 2)
    assert rel_err(blur.A(img), R.blur_A(img.contiguous(), torch.ones(1, 1, 3, 3) / 9, "circular")) < 1e-6


def test_shape_and_batch_mismatches_raise_the_reference_error_types():
    import deepinv_b200 as dinv

    p = dinv.physics.MRI(mask=torch.ones(3, 2, 8, 8), img_size=(2, 8, 8))
    with pytest.raises(ValueError):
        p.A(torch.zeros(2, 2, 8, 8))                       # mask batch 3 vs input batch 2
    with pytest.raises(ValueError):
        p.A(torch.zeros(3, 2, 8, 9))                       # wrong image size
    t = dinv.physics.Tomography(angles=4, img_width=8, normalize=False)
    with pytest.raises(ValueError):
        t.A(torch.zeros(1, 1, 9, 9))                       # tomography.py:244-247
    with pytest.raises(ValueError):
        dinv.physics.Blur(filter=torch.ones(1, 1, 3, 3), padding="bogus").A(torch.zeros(1, 1, 8, 8))  # convolution.py:24-39
    with pytest.raises(AssertionError):
        dinv.physics.Blur(filter=torch.ones(1, 2, 3, 3), padding="valid").A(torch.zeros(1, 3, 8, 8))  # convolution.py:774-784
    with pytest.raises(ValueError):
        dinv.physics.MultiCoilMRI(coil_maps=torch.ones(1, 2, 8, 8), img_size=(2, 8, 8))              # real-valued maps

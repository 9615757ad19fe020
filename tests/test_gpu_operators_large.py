"""GPU parity at sizes beyond the fixtures (oracle evaluated on CPU in the same process) and size-independent
properties at the BASELINE.json sizes."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_tomography_vs_oracle_64(dev):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    W, nang = 64, 30
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 1, W, W, generator=gen)
    ang = R.default_angles(nang)
    phys = dinv.physics.Tomography(angles=nang, img_width=W, normalize=False, device=dev)
    y = R.radon_forward(x, ang)
    assert rel_err(phys.A(x.to(dev)), y) < TOL
    v = torch.randn(*y.shape, generator=gen)
    assert rel_err(phys.A_adjoint(v.to(dev)), R.radon_adjoint(v, ang, W)) < TOL
    assert rel_err(phys.A_dagger(y.to(dev), fbp=True), R.tomography_fbp(y, ang, W)) < TOL


def test_tomography_cfg3_properties(dev):
    """512^2, 180 angles (P = 725): shapes, view layout, adjointness of the exact transpose, FBP sanity"""
    import deepinv_b200 as dinv

    phys = dinv.physics.Tomography(angles=180, img_width=512, normalize=False, device=dev)
    assert phys.P == 725
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(2, 1, 512, 512, device=dev, generator=gen)
    y = phys.A(x)
    assert y.shape == (2, 1, 725, 180) and y.stride()[-2:] == (1, 725)
    v = torch.randn(2, 1, 725, 180, device=dev, generator=gen)
    lhs, rhs = (y.double() * v.double()).sum(), (x.double() * phys.A_adjoint(v).double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-4
    disc = torch.zeros(1, 1, 512, 512, device=dev)
    disc[..., 156:356, 156:356] = 1.0
    rec = phys.A_dagger(phys.A(disc), fbp=True)
    assert float((rec - disc).abs().mean()) < 0.05  # FBP reconstructs a centred square


def test_blur_31x31_vs_oracle(dev):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    gen = torch.Generator().manual_seed(0)
    x = torch.rand(2, 1, 96, 160, generator=gen)
    f = torch.rand(1, 1, 31, 31, generator=gen)
    f /= f.sum()
    for pad in ("valid", "circular", "replicate", "reflect", "constant"):
        phys = dinv.physics.Blur(filter=f.to(dev), padding=pad, device=dev)
        y = R.blur_A(x, f, pad)
        assert rel_err(phys.A(x.to(dev)), y) < TOL
        v = torch.rand(*y.shape, generator=gen)
        assert rel_err(phys.A_adjoint(v.to(dev)), R.blur_At(v, f, pad, 96, 160)) < TOL


def test_blur_cfg5_properties(dev):
    """1024^2 with a 31x31 PSF: Blur(circular) == BlurFFT, adjointness for every padding"""
    import deepinv_b200 as dinv

    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(2, 1, 1024, 1024, device=dev, generator=gen)
    f = torch.rand(1, 1, 31, 31, device=dev, generator=gen)
    f /= f.sum()
    fft = dinv.physics.BlurFFT(img_size=(1, 1024, 1024), filter=f, device=dev)
    circ = dinv.physics.Blur(filter=f, padding="circular", device=dev)
    assert rel_err(circ.A(x), fft.A(x)) < TOL
    assert rel_err(circ.A_adjoint(x), fft.A_adjoint(x)) < TOL
    for pad in ("valid", "circular", "replicate", "reflect", "constant"):
        phys = dinv.physics.Blur(filter=f, padding=pad, device=dev)
        y = phys.A(x)
        v = torch.rand_like(y)
        lhs, rhs = (y.double() * v.double()).sum(), (x.double() * phys.A_adjoint(v).double()).sum()
        assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-5


def test_multicoil_cfg4_properties(dev):
    import deepinv_b200 as dinv

    B, N, H, W = 4, 8, 320, 320
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, 2, H, W, device=dev, generator=gen)
    maps = torch.view_as_complex(torch.randn(1, N, H, W, 2, device=dev, generator=gen))
    maps = maps / maps.abs().pow(2).sum(1, keepdim=True).sqrt()
    mask = (torch.rand(1, 1, 1, W, device=dev, generator=gen) > 0.8).float().expand(1, 2, H, W).contiguous()
    phys = dinv.physics.MultiCoilMRI(mask=mask, coil_maps=maps, img_size=(2, H, W), device=dev)
    y = phys.A(x)
    assert y.shape == (B, 2, N, H, W)
    v = torch.randn_like(y)
    lhs, rhs = (y.double() * v.double()).sum(), (x.double() * phys.A_adjoint(v).double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-5
    full = dinv.physics.MultiCoilMRI(mask=torch.ones(H, W, device=dev), coil_maps=maps, img_size=(2, H, W), device=dev)
    assert rel_err(full.A_adjoint(full.A(x)), x) < 1e-5  # sum |S_n|^2 = 1 and F unitary

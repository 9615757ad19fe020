"""GPU parity: MRI operators, fp32 denoisers and the PnP loops against the reference's golden vectors
and the oracle.  Tolerance: 1e-5 relative L2 in fp32 (BASELINE.json north star); exact zero pattern
for masked k-space."""
import pytest
import torch

from conftest import golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _cu(g, dev):
    out = {}
    for k, v in g.items():
        out[k] = {kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else v.to(dev)
    return out


@pytest.mark.parametrize("shape", [(4, 64, 64), (2, 256, 256), (2, 320, 320), (3, 48, 80), (2, 37, 31), (1, 128, 96)])
def test_mri_vs_oracle(shape, dev):
    """sizes beyond the fixtures: smooth (radix 2/3/5) and prime-factor sizes, per-sample line masks"""
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    B, H, W = shape
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, 2, H, W, generator=gen)
    cols = (torch.rand(B, 1, 1, W, generator=gen) > 0.6).float().expand(B, 2, H, W).contiguous()
    full = (torch.rand(B, 2, H, W, generator=gen) > 0.5).float()
    for mask in (cols, full, full[:1]):
        phys = dinv.physics.MRI(mask=mask.to(dev), img_size=(2, H, W), device=dev)
        y = R.mri_A(x, mask)
        assert rel_err(phys.A(x.to(dev)), y) < TOL
        assert rel_err(phys.A_adjoint(y.to(dev)), R.mri_At(y, mask)) < TOL
        assert rel_err(phys.A_adjoint_A(x.to(dev)), R.mri_AtA(x, mask)) < TOL
        assert rel_err(phys.prox_l2(x.to(dev), y.to(dev), 1.3), R.mri_prox_l2(x, y, mask, 1.3)) < TOL
        u = torch.randn(B, 2, H, W, generator=gen).to(dev)
        assert abs(float(phys.adjointness_test(u))) < 1e-3 * B * H * W ** 0.5


def test_mri_full_size_properties(dev):
    """cfg2 size (64 x 256^2): adjointness, unitarity of V and projection property at the benchmark size"""
    import deepinv_b200 as dinv
    from deepinv_b200 import ops

    B, H, W = 64, 256, 256
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, 2, H, W, device=dev, generator=gen)
    mask = (torch.rand(B, 1, 1, W, device=dev, generator=gen) > 0.75).float().expand(B, 2, H, W).contiguous()
    phys = dinv.physics.MRI(mask=mask, img_size=(2, H, W), device=dev)
    v = torch.randn(B, 2, H, W, device=dev, generator=gen)
    Ax, Atv = phys.A(x), phys.A_adjoint(v)
    lhs = ops.batched_dot(Ax, v).double().sum()
    rhs = ops.batched_dot(x, Atv).double().sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-5
    assert rel_err(phys.V(phys.V_adjoint(x)), x) < 1e-6 * 5
    assert rel_err(phys.A_adjoint_A(phys.A_adjoint_A(x)), phys.A_adjoint_A(x)) < 1e-5  # (A^T A)^2 = A^T A for 0/1 masks
    assert torch.equal(Ax == 0, mask == 0) or float(((Ax == 0) != (mask == 0)).float().mean()) < 1e-6


def test_elementwise_and_dots(dev):
    from deepinv_b200 import ops

    gen = torch.Generator().manual_seed(0)
    for n in (1, 7, 4099, 1 << 20):
        x, y, z = (torch.randn(n, generator=gen) for _ in range(3))
        out = ops.axpbypcz(x.to(dev), 0.5, y.to(dev), -2.0, z.to(dev), 3.0)
        assert rel_err(out, 0.5 * x - 2.0 * y + 3.0 * z) < 1e-6
    a, b = torch.randn(5, 3, 37, 41, generator=gen), torch.randn(5, 3, 37, 41, generator=gen)
    d = ops.batched_dot(a.to(dev), b.to(dev))
    assert rel_err(d, (a.double() * b.double()).flatten(1).sum(1)) < 1e-6
    s = torch.rand(5, generator=gen)
    assert rel_err(ops.batched_axpy(a.to(dev), b.to(dev), s.to(dev), -1.0), a - s.view(-1, 1, 1, 1) * b) < 1e-6


def test_conv_f32_vs_torch_cpu(dev):
    """fp32 conv kernels against the oracle's ATen CPU convolutions at DRUNet channel counts"""
    import torch.nn.functional as F

    from deepinv_b200 import ops

    gen = torch.Generator().manual_seed(1)
    for (B, Cin, Cout, H, W) in [(2, 3, 64, 40, 72), (1, 64, 64, 33, 17), (2, 128, 128, 16, 16), (1, 64, 2, 32, 32)]:
        x, xa = torch.randn(B, Cin, H, W, generator=gen), torch.randn(B, Cin, H, W, generator=gen)
        w = torch.randn(Cout, Cin, 3, 3, generator=gen) / (3 * Cin ** 0.5)
        b, r = torch.randn(Cout, generator=gen), torch.randn(B, Cout, H, W, generator=gen)
        out = ops.conv_f32(x.to(dev), w.to(dev), bias=b.to(dev), xadd=xa.to(dev), res=r.to(dev), relu=True)
        assert rel_err(out, F.relu(F.conv2d(x + xa, w, b, padding=1)) + r) < TOL
    x = torch.randn(2, 64, 16, 24, generator=gen)
    w = torch.randn(128, 64, 2, 2, generator=gen) / 16
    assert rel_err(ops.conv_f32(x.to(dev), w.to(dev), kind=1), F.conv2d(x, w, stride=2)) < TOL
    wt = torch.randn(64, 32, 2, 2, generator=gen) / 8
    xa = torch.randn(2, 64, 16, 24, generator=gen)
    assert rel_err(ops.conv_f32(x.to(dev), wt.to(dev), kind=2, xadd=xa.to(dev)), F.conv_transpose2d(x + xa, wt, stride=2)) < TOL

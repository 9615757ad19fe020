"""Kernel-logic tests without a GPU: the SIMT kernel sources of libdinvk are compiled for the host
against tests/emul/cuda_emul.h (one CUDA thread = one host thread, real barriers) and compared with
the oracle.  This checks index math, tile geometry, barriers and the C-ABI argument handling; it says
nothing about performance and is not a product path (see tests/emul/cuda_emul.h)."""
import ctypes as C

import pytest
import torch

from conftest import load_golden, rel_err
from emul_util import emul_lib, ptr, spectral_emul
from oracle import ref_ops as R

TOL = 1e-5
G_NONE, G_MASK, G_SQ, G_INV, G_PINV, G_CMUL, G_CMULC = range(7)


def full_strides(mask, H, W):
    return (2 * H * W if mask.shape[0] > 1 else 0, H * W, W)


@pytest.mark.parametrize("name", ["mri_16x12_full", "mri_32x32_lines", "mri_17x11_odd", "mri_20x24_shared"])
def test_spectral_mri_golden(name):
    g = load_golden(name)
    x, m, y, z, gam = g["x"], g["mask"].contiguous(), g["y"], g["z"], float(g["gamma"])
    H, W = x.shape[-2:]
    st = full_strides(m, H, W)
    yk = spectral_emul(x, H, W, True, False, G_MASK, m, st)
    assert rel_err(yk, y) < TOL and torch.equal(yk == 0, y == 0)
    assert rel_err(spectral_emul(y, H, W, False, True, G_MASK, m, st), g["At"]) < TOL
    assert rel_err(spectral_emul(x, H, W, True, True, G_SQ, m, st), g["AtA"]) < TOL
    assert rel_err(spectral_emul(y, H, W, False, False, G_SQ, m, st), g["AAt"]) < TOL
    assert rel_err(spectral_emul(y, H, W, False, True, G_PINV, m, st), g["dagger"]) < TOL
    assert rel_err(spectral_emul(g["At"], H, W, True, True, G_INV, m, st, p1=z, a1=1 / gam, c=1 / gam), g["prox"]) < TOL
    assert rel_err(spectral_emul(x, H, W, True, False), g["Vt"]) < TOL
    assert rel_err(spectral_emul(x, H, W, False, True), g["V"]) < TOL


@pytest.mark.parametrize("shape", [(2, 256, 256), (1, 320, 320), (2, 48, 80), (3, 8, 5), (2, 1, 16), (1, 64, 1024)])
def test_spectral_sizes_vs_oracle(shape):
    B, H, W = shape
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, 2, H, W, generator=gen)
    full = (torch.rand(B, 2, H, W, generator=gen) > 0.5).float()
    cols1 = (torch.rand(B, 1, 1, W, generator=gen) > 0.5).float()          # complex-scalar line mask (B,1,1,W)
    colsx = cols1.expand(B, 2, H, W).contiguous()
    assert rel_err(spectral_emul(x, H, W, True, False, G_MASK, full, full_strides(full, H, W)), R.mri_A(x, full)) < TOL
    assert rel_err(spectral_emul(x, H, W, False, True, G_MASK, full, full_strides(full, H, W)), R.mri_At(x, full)) < TOL
    # fused single-pass path (line mask) == three-pass path == oracle
    st1 = (W if B > 1 else 0, 0, 0)
    aty = R.mri_At(R.mri_A(x, colsx), colsx)
    step = spectral_emul(x, H, W, True, True, G_SQ, cols1.contiguous(), st1, e0=-0.9, q0=x, e1=1.0, q1=aty, e2=0.9)
    ref = x - 0.9 * (R.mri_AtA(x, colsx) - aty)
    assert rel_err(step, ref) < TOL
    step3 = spectral_emul(x, H, W, True, True, G_SQ, colsx, full_strides(colsx, H, W), e0=-0.9, q0=x, e1=1.0, q1=aty, e2=0.9)
    assert rel_err(step3, ref) < TOL


def test_spectral_multicoil_golden():
    g = load_golden("mcmri_16x20")
    maps = torch.complex(g["maps_re"], g["maps_im"]).contiguous()
    x, m, y = g["x"], g["mask"].contiguous(), g["y"].contiguous()
    B, N, H, W = maps.shape
    st = full_strides(m, H, W)
    yk = spectral_emul(x, H, W, True, False, G_MASK, m, st, ncoil=N, coil_mode=1, coil_maps=maps, out_shape=(B, 2, N, H, W))
    assert rel_err(yk, y) < TOL
    xa = spectral_emul(y, H, W, False, True, G_MASK, m, st, ncoil=N, coil_mode=2, coil_maps=maps, out_shape=(B, 2, H, W))
    assert rel_err(xa, g["At"]) < TOL
    xr = spectral_emul(y, H, W, False, True, G_MASK, m, st, ncoil=N, coil_mode=3, coil_maps=maps, out_shape=(B, 1, H, W))
    assert rel_err(xr, g["At_rss"]) < TOL


def test_conv_f32_golden_drunet_layers():
    """the fp32 conv kernels reproduce a reference DRUNet forward when chained like models/drunet.py"""
    import torch.nn.functional as F

    lib = emul_lib()
    gen = torch.Generator().manual_seed(0)

    def conv(x, w, kind=0, bias=None, xadd=None, res=None, relu=False):
        B, Cin, H, W = x.shape
        Cout = w.shape[1] if kind == 2 else w.shape[0]
        shape = (B, Cout, 2 * H, 2 * W) if kind == 2 else (B, Cout, H // 2, W // 2) if kind == 1 else (B, Cout, H, W)
        out = torch.empty(shape)
        rc = lib.dinvk_conv_f32(ptr(x.contiguous()), ptr(xadd), ptr(w.contiguous()), ptr(bias), ptr(res), ptr(out), B, Cin, Cout,
                                H, W, kind, int(relu), None)
        assert rc == 0, lib.dinvk_last_error()
        return out

    for (B, Cin, Cout, H, W) in [(2, 3, 64, 16, 40), (1, 64, 64, 9, 33), (2, 20, 37, 8, 8), (1, 64, 2, 16, 32)]:
        x, xa = torch.randn(B, Cin, H, W, generator=gen), torch.randn(B, Cin, H, W, generator=gen)
        w, b = torch.randn(Cout, Cin, 3, 3, generator=gen) * 0.1, torch.randn(Cout, generator=gen)
        r = torch.randn(B, Cout, H, W, generator=gen)
        assert rel_err(conv(x, w, 0, b, xa, r, True), F.relu(F.conv2d(x + xa, w, b, padding=1)) + r) < TOL
    x = torch.randn(2, 20, 6, 10, generator=gen)
    w = torch.randn(7, 20, 2, 2, generator=gen) * 0.1
    assert rel_err(conv(x, w, 1), F.conv2d(x, w, stride=2)) < TOL
    wt = torch.randn(20, 5, 2, 2, generator=gen) * 0.1
    assert rel_err(conv(x, wt, 2, xadd=x), F.conv_transpose2d(2 * x, wt, stride=2)) < TOL


def test_elementwise_cg_helpers():
    lib = emul_lib()
    gen = torch.Generator().manual_seed(0)
    for n in (1, 5, 1027, 8192):
        x, y, z = (torch.randn(n, generator=gen) for _ in range(3))
        out = torch.empty(n)
        assert lib.dinvk_axpbypcz(ptr(out), ptr(x), 0.5, ptr(y), -2.0, ptr(z), 3.0, n, None) == 0
        assert rel_err(out, 0.5 * x - 2 * y + 3 * z) < 1e-6
        assert lib.dinvk_axpbypcz(ptr(out), ptr(x), 2.0, None, 0.0, None, 0.0, n, None) == 0
        assert rel_err(out, 2 * x) < 1e-7
    B, n = 3, 5000
    a, b = torch.randn(B, n, generator=gen), torch.randn(B, n, generator=gen)
    d = torch.empty(B)
    nb = lib.dinvk_batched_dot_workspace_bytes(B, n)
    ws = torch.zeros(nb, dtype=torch.uint8)
    assert lib.dinvk_batched_dot(ptr(d), ptr(a), ptr(b), B, n, ptr(ws), nb, None) == 0
    assert rel_err(d, (a.double() * b.double()).sum(1)) < 1e-6
    s = torch.rand(B, generator=gen)
    out = torch.empty(B, n)
    assert lib.dinvk_batched_axpy(ptr(out), ptr(a), ptr(b), ptr(s), -1.0, B, n, None) == 0
    assert rel_err(out, a - s[:, None] * b) < 1e-6
    # CG scalars: alpha, beta + sticky convergence flag
    num, den, bn = torch.tensor([1.0, 2.0, 3.0]), torch.tensor([2.0, 4.0, 8.0]), torch.tensor([1.0, 0.0, 4.0])
    flag = torch.zeros(1, dtype=torch.int32)
    o = torch.empty(3)
    assert lib.dinvk_cg_scalars(0, ptr(o), ptr(num), ptr(den), 1e-8, None, 0.0, ptr(flag), 3, None) == 0
    assert torch.allclose(o, num / (den + 1e-8))
    assert lib.dinvk_cg_scalars(1, ptr(o), ptr(num), ptr(den), 1e-8, ptr(bn), 0.5, ptr(flag), 3, None) == 0
    assert int(flag) == 0  # 1 < 0.5*1 false
    assert lib.dinvk_cg_scalars(1, ptr(o), ptr(num), ptr(den), 1e-8, ptr(bn), 4.0, ptr(flag), 3, None) == 0
    assert int(flag) == 1  # 1<4, 2<4 (bn=0 -> 1), 3<16
    assert lib.dinvk_cg_scalars(0, ptr(o), ptr(num), ptr(den), 1e-8, None, 0.0, ptr(flag), 3, None) == 0
    assert torch.equal(o, torch.zeros(3))  # frozen after convergence


def test_ddrm_update_vs_oracle_step():
    """one DDRM spectral update (sampling/diffusion.py:196-217) on the three mask cases"""
    import math

    lib = emul_lib()
    gen = torch.Generator().manual_seed(0)
    n, mask_n = 2 * 2 * 8 * 8, 2 * 8 * 8
    mask = (torch.rand(mask_n, generator=gen) > 0.4).float() * torch.rand(mask_n, generator=gen)
    xb, xp, yb, noise = (torch.randn(n, generator=gen) for _ in range(4))
    sig_t, sig_prev, sn, eta, etab, eps = 0.5, 0.75, 0.3, 0.85, 1.0, 1e-6
    c = math.sqrt(1 - eta ** 2)
    out = torch.empty(n)
    assert lib.dinvk_ddrm_update(ptr(out), ptr(xb), ptr(xp), ptr(yb), ptr(mask), ptr(noise), n, mask_n, sig_t, sig_prev, sn,
                                 eta, etab, c * sig_t, eps, 0, None) == 0
    m = mask.repeat(2)
    case = m > sn
    nsr = torch.where(case, sn / (m + eps), torch.zeros_like(m))
    case2, case3 = case & (sig_t < nsr), case & (sig_t >= nsr)
    mean = xb + c * sig_t * (xp - xb) / sig_prev
    mean[case2] = xb[case2] + c * sig_t * (yb[case2] - xb[case2]) / (nsr[case2] + eps)
    mean[case3] = (1 - etab) * xb[case3] + etab * yb[case3]
    std = torch.ones(n) * eta * sig_t
    std[case3] = (sig_t ** 2 - (nsr[case3] * etab) ** 2).clamp(min=0).sqrt()
    assert rel_err(out, mean + std * noise / math.sqrt(2.0)) < 1e-6
    assert case2.any() and case3.any() and (~case).any()

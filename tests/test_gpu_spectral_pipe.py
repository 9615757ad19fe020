"""GPU parity of the persistent bulk-copy-pipelined 256x256 spectral kernels (csrc/spectral_pipe.cuh).

Three checkers: the oracle (CPU restatement of the reference, small batch), the older tile-pass kernels
(`DINVK_NO_PIPE_FFT=1`, same library, every operand combination at a batch large enough that each persistent
CTA walks several tiles of its ring), and size-independent properties at the cfg2 batch.
Tolerance 1e-5 relative L2 (north star); the two kernel families agree to ~1e-6 (same butterflies, different order)."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

H = W = 256


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


class tile_passes:
    """run the enclosed calls through the older tile-pass kernels"""

    def __enter__(self):
        os.environ["DINVK_NO_PIPE_FFT"] = "1"

    def __exit__(self, *exc):
        os.environ.pop("DINVK_NO_PIPE_FFT", None)


def _masks(B, gen):
    cols = (torch.rand(B, 1, 1, W, generator=gen) > 0.6).float().expand(B, 2, H, W).contiguous()
    full = (torch.rand(B, 2, H, W, generator=gen) > 0.5).float()
    return {"line": cols, "full": full, "shared_full": full[:1].contiguous(), "shared_line": cols[:1].contiguous()}


@pytest.mark.parametrize("B", [1, 5])
def test_pipe_vs_oracle(B, dev):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, 2, H, W, generator=gen)
    for name, mask in _masks(B, gen).items():
        phys = dinv.physics.MRI(mask=mask.to(dev), img_size=(2, H, W), device=dev)
        y = R.mri_A(x, mask)
        got = phys.A(x.to(dev))
        assert rel_err(got, y) < 1e-5, name
        assert torch.equal(got.cpu() == 0, y == 0) or float(((got.cpu() == 0) != (y == 0)).float().mean()) < 1e-6
        assert rel_err(phys.A_adjoint(y.to(dev)), R.mri_At(y, mask)) < 1e-5, name
        assert rel_err(phys.A_adjoint_A(x.to(dev)), R.mri_AtA(x, mask)) < 1e-5, name
        assert rel_err(phys.prox_l2(x.to(dev), y.to(dev), 0.7), R.mri_prox_l2(x, y, mask, 0.7)) < 1e-5, name
        assert rel_err(phys.A_dagger(y.to(dev)), R.mri_At(y, mask)) < 1e-5, name  # 0/1 mask: pinv == adjoint
        aty = R.mri_At(y, mask)
        want = x - 0.9 * (R.mri_AtA(x, mask) - aty)
        assert rel_err(phys.normal_step(x.to(dev), aty.to(dev), 0.9), want) < 1e-5, name


def test_pipe_vs_tile_passes_many_tiles(dev):
    """B = 45 -> 720 tiles over 296 persistent CTAs: every CTA refills its ring; all operand combinations"""
    from deepinv_b200 import _ffi, ops

    B = 45
    gen = torch.Generator(device=dev).manual_seed(2)
    r = lambda: torch.randn(B, 2, H, W, device=dev, generator=gen)
    x, p1, q0, q1 = r(), r(), r(), r()
    line = ops.mask_spec_from_real((torch.rand(B, 1, 1, W, device=dev, generator=gen) > 0.7).float().expand(B, 2, H, W).contiguous(), H, W)
    full = ops.mask_spec_from_real(torch.rand(B, 2, H, W, device=dev, generator=gen), H, W)
    shared = ops.mask_spec_from_real(torch.rand(1, 2, H, W, device=dev, generator=gen), H, W)
    assert line.sh == 0 and full.sh == W
    cb = torch.rand(B, device=dev, generator=gen) + 0.5
    cases = []
    for mname, m in (("line", line), ("full", full), ("shared", shared)):
        for gmode in (_ffi.G_MASK, _ffi.G_SQ, _ffi.G_INV_SQ_PLUS_C, _ffi.G_PINV):
            cases.append((f"A {mname} g{gmode}", dict(fwd=True, inv=False, gmode=gmode, mask=m, c=0.8)))
            cases.append((f"At {mname} g{gmode}", dict(fwd=False, inv=True, gmode=gmode, mask=m, c=0.8)))
        cases.append((f"A {mname} epilogue", dict(fwd=True, inv=False, gmode=_ffi.G_MASK, mask=m, a0=0.5, p1=p1, a1=-1.5, e0=2.0,
                                                   q0=q0, e1=0.25, q1=q1, e2=-0.75)))
        cases.append((f"At {mname} epilogue", dict(fwd=False, inv=True, gmode=_ffi.G_MASK, mask=m, a0=0.5, p1=p1, a1=-1.5, e0=2.0,
                                                    q0=q0, e1=0.25, q1=q1, e2=-0.75)))
        cases.append((f"At {mname} c_batch", dict(fwd=False, inv=True, gmode=_ffi.G_INV_SQ_PLUS_C, mask=m, c_batch=cb)))
    for gmode in (_ffi.G_MASK, _ffi.G_SQ, _ffi.G_INV_SQ_PLUS_C, _ffi.G_PINV):
        cases.append((f"fused line g{gmode}", dict(fwd=True, inv=True, gmode=gmode, mask=line, c=1.3)))
    cases += [
        ("fused none", dict(fwd=True, inv=True)),
        ("A none", dict(fwd=True, inv=False)),
        ("At none", dict(fwd=False, inv=True)),
        ("A uncentred", dict(fwd=True, inv=False, centered=False, gmode=_ffi.G_MASK, mask=full)),
        ("At uncentred", dict(fwd=False, inv=True, centered=False, gmode=_ffi.G_MASK, mask=full)),
        ("fused uncentred", dict(fwd=True, inv=True, centered=False, gmode=_ffi.G_SQ, mask=line)),
        ("normal step", dict(fwd=True, inv=True, gmode=_ffi.G_SQ, mask=line, e0=-0.9, q0=x, e1=1.0, q1=q1, e2=0.9)),
        ("fused q0 != p0", dict(fwd=True, inv=True, gmode=_ffi.G_SQ, mask=line, e0=-0.9, q0=q0, e1=1.0, q1=q1, e2=0.9)),
        ("fused a0", dict(fwd=True, inv=True, gmode=_ffi.G_SQ, mask=line, a0=0.3, e0=-0.9, q0=x, e1=1.0)),
        ("prox", dict(fwd=True, inv=True, gmode=_ffi.G_INV_SQ_PLUS_C, mask=line, p1=p1, a1=1.0 / 0.7, c=1.0 / 0.7)),
        ("prox c_batch", dict(fwd=True, inv=True, gmode=_ffi.G_INV_SQ_PLUS_C, mask=line, p1=p1, a1=1.0, c_batch=cb)),
    ]
    for name, kw in cases:
        got = ops.spectral(x, H, W, **kw)
        with tile_passes():
            want = ops.spectral(x, H, W, **kw)
        assert rel_err(got, want) < 3e-6, name
    torch.cuda.synchronize()


def test_pipe_properties_cfg2(dev):
    """64 x 256^2: adjointness <Ax, v> = <x, A^T v>, unitarity, idempotence of A^T A for 0/1 line masks, exact zeros"""
    import deepinv_b200 as dinv
    from deepinv_b200 import ops

    B = 64
    gen = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(B, 2, H, W, device=dev, generator=gen)
    v = torch.randn(B, 2, H, W, device=dev, generator=gen)
    mask = (torch.rand(B, 1, 1, W, device=dev, generator=gen) > 0.75).float().expand(B, 2, H, W).contiguous()
    phys = dinv.physics.MRI(mask=mask, img_size=(2, H, W), device=dev)
    Ax, Atv = phys.A(x), phys.A_adjoint(v)
    lhs = ops.batched_dot(Ax, v).double().sum()
    rhs = ops.batched_dot(x, Atv).double().sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-5
    assert rel_err(phys.V(phys.V_adjoint(x)), x) < 5e-6
    AtA = phys.A_adjoint_A(x)
    assert rel_err(phys.A_adjoint(phys.A(x)), AtA) < 5e-6          # two-pass A, A^T against the fused row pass
    assert rel_err(phys.A_adjoint_A(AtA), AtA) < 1e-5
    assert torch.equal(Ax == 0, mask == 0) or float(((Ax == 0) != (mask == 0)).float().mean()) < 1e-6
    # Parseval: ||F x|| = ||x|| for the orthonormal transform
    nx = float(ops.batched_dot(x, x).double().sum())
    Vx = phys.V_adjoint(x)
    assert abs(float(ops.batched_dot(Vx, Vx).double().sum()) - nx) / nx < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# 320 x 320 (cfg4): two-pass kernels of csrc/spectral_pipe320.cuh
# ---------------------------------------------------------------------------------------------------------------------
def test_pipe320_vs_oracle(dev):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    B, Hh, Ww, N = 3, 320, 320, 4
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(B, 2, Hh, Ww, generator=gen)
    cols = (torch.rand(B, 1, 1, Ww, generator=gen) > 0.8).float().expand(B, 2, Hh, Ww).contiguous()
    full = (torch.rand(B, 2, Hh, Ww, generator=gen) > 0.5).float()
    for mask in (cols, full, full[:1].contiguous()):
        phys = dinv.physics.MRI(mask=mask.to(dev), img_size=(2, Hh, Ww), device=dev)
        y = R.mri_A(x, mask)
        assert rel_err(phys.A(x.to(dev)), y) < 1e-5
        assert rel_err(phys.A_adjoint(y.to(dev)), R.mri_At(y, mask)) < 1e-5
    maps = torch.view_as_complex(torch.randn(1, N, Hh, Ww, 2, generator=gen).contiguous())
    maps = maps / maps.abs().pow(2).sum(1, keepdim=True).sqrt()
    mphys = dinv.physics.MultiCoilMRI(mask=cols.to(dev), coil_maps=maps.to(dev), img_size=(2, Hh, Ww), device=dev)
    y = R.mcmri_A(x, cols, maps)
    assert rel_err(mphys.A(x.to(dev)), y) < 1e-5
    assert rel_err(mphys.A_adjoint(y.to(dev)), R.mcmri_At(y, cols, maps)) < 1e-5
    assert rel_err(mphys.A_adjoint(y.to(dev), rss=True), R.mcmri_At(y, cols, maps, use_rss=True)) < 1e-5


def test_pipe320_vs_tile_passes_many_tiles(dev):
    from deepinv_b200 import _ffi, ops

    Hh = Ww = 320
    B, N = 24, 8
    gen = torch.Generator(device=dev).manual_seed(6)
    r = lambda *shape: torch.randn(*shape, device=dev, generator=gen)
    x, p1, q0, q1 = r(B, 2, Hh, Ww), r(B, 2, Hh, Ww), r(B, 2, Hh, Ww), r(B, 2, Hh, Ww)
    line = ops.mask_spec_from_real((torch.rand(B, 1, 1, Ww, device=dev, generator=gen) > 0.7).float().expand(B, 2, Hh, Ww).contiguous(), Hh, Ww)
    full = ops.mask_spec_from_real(torch.rand(B, 2, Hh, Ww, device=dev, generator=gen), Hh, Ww)
    shared = ops.mask_spec_from_real(torch.rand(1, 2, Hh, Ww, device=dev, generator=gen), Hh, Ww)
    cb = torch.rand(B, device=dev, generator=gen) + 0.5
    cases = [("A none", dict(fwd=True, inv=False)), ("At none", dict(fwd=False, inv=True)),
             ("A uncentred", dict(fwd=True, inv=False, centered=False, gmode=_ffi.G_MASK, mask=full)),
             ("At uncentred", dict(fwd=False, inv=True, centered=False, gmode=_ffi.G_MASK, mask=full))]
    for mname, m in (("line", line), ("full", full), ("shared", shared)):
        for gmode in (_ffi.G_MASK, _ffi.G_SQ, _ffi.G_INV_SQ_PLUS_C, _ffi.G_PINV):
            cases.append((f"A {mname} g{gmode}", dict(fwd=True, inv=False, gmode=gmode, mask=m, c=0.8)))
            cases.append((f"At {mname} g{gmode}", dict(fwd=False, inv=True, gmode=gmode, mask=m, c=0.8)))
        ep = dict(gmode=_ffi.G_MASK, mask=m, a0=0.5, p1=p1, a1=-1.5, e0=2.0, q0=q0, e1=0.25, q1=q1, e2=-0.75)
        cases.append((f"A {mname} epilogue", dict(fwd=True, inv=False, **ep)))
        cases.append((f"At {mname} epilogue", dict(fwd=False, inv=True, **ep)))
        cases.append((f"At {mname} c_batch", dict(fwd=False, inv=True, gmode=_ffi.G_INV_SQ_PLUS_C, mask=m, c_batch=cb)))
    for name, kw in cases:
        got = ops.spectral(x, Hh, Ww, **kw)
        with tile_passes():
            want = ops.spectral(x, Hh, Ww, **kw)
        assert rel_err(got, want) < 3e-6, name
    # multi-coil: batch 5 x 8 coils = 40 coil images (800 pass-1 tiles); per-sample and shared maps, per-sample line masks
    Bm = 5
    xm = r(Bm, 2, Hh, Ww)
    for shared_maps in (True, False):
        maps = torch.view_as_complex(r(1 if shared_maps else Bm, N, Hh, Ww, 2))
        lm = ops.mask_spec_from_real((torch.rand(Bm, 1, 1, Ww, device=dev, generator=gen) > 0.8).float().expand(Bm, 2, Hh, Ww).contiguous(), Hh, Ww)
        kwA = dict(fwd=True, inv=False, gmode=_ffi.G_MASK, mask=lm, ncoil=N, coil_mode=1, coil_maps=maps)
        y = ops.spectral(xm, Hh, Ww, **kwA)
        with tile_passes():
            y0 = ops.spectral(xm, Hh, Ww, **kwA)
        assert y.shape == (Bm, 2, N, Hh, Ww) and rel_err(y, y0) < 3e-6
        for mode in (2, 3):
            kwT = dict(fwd=False, inv=True, gmode=_ffi.G_MASK, mask=lm, ncoil=N, coil_mode=mode, coil_maps=maps, e0=(0.7 if mode == 2 else 1.0))
            v = ops.spectral(y0, Hh, Ww, **kwT)
            with tile_passes():
                v0 = ops.spectral(y0, Hh, Ww, **kwT)
            assert rel_err(v, v0) < 3e-6, (shared_maps, mode)
    torch.cuda.synchronize()

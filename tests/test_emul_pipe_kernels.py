"""The benchmark-size spectral kernels — the persistent, bulk-copy-pipelined 256x256 (cfg2) and 320x320 (cfg4) kernels of
csrc/spectral_pipe*.cuh — executed on the host: tests/emul models the mbarrier / cp.async.bulk primitives (a phase completes
when all expected arrivals AND bytes are in), so tile schedule, ring indexing, barrier parities, butterflies, twiddles and
epilogues of the very kernels bench.py times are checked against the oracle in the GPU-less container.  (The same cases run
on the device in tests/test_gpu_spectral_pipe.py.)"""
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
    yield lib
    ops._ws_cache.clear()


def _masks(kind, B, H, W, gen):
    if kind == "lines":
        return (torch.rand(B, 1, 1, W, generator=gen) > 0.75).float().expand(B, 2, H, W).contiguous()
    if kind == "shared":
        return (torch.rand(1, 1, H, W, generator=gen) > 0.5).float().expand(1, 2, H, W).contiguous()
    return (torch.rand(B, 2, H, W, generator=gen) > 0.5).float()


@pytest.mark.parametrize("kind", ["lines", "full"])
def test_pipe256_operators(kind, emul_backend, monkeypatch):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    gen = torch.Generator().manual_seed(7)
    B, H, W = 3, 256, 256  # 48 tiles: several tiles per "CTA" ring in the emulated persistent grid
    x = torch.randn(B, 2, H, W, generator=gen)
    z = torch.randn(B, 2, H, W, generator=gen)
    mask = _masks(kind, B, H, W, gen)
    p = dinv.physics.MRI(mask=mask, img_size=(2, H, W))
    n0 = emul_backend.dinvk_launch_count()
    y = p.A(x)
    assert emul_backend.dinvk_launch_count() - n0 == 2  # pass 1 + pass 2 of the pipelined path (tile passes: also 2)
    ref_y = R.mri_A(x, mask)
    assert rel_err(y, ref_y) < 1e-6 and torch.equal(y == 0, ref_y == 0)
    aty = R.mri_At(ref_y, mask)
    assert rel_err(p.A_adjoint(ref_y), aty) < 1e-6
    if kind != "lines":  # fused passes with h-dependent masks take the general tile passes (tests/test_emul_kernels.py)
        return
    assert rel_err(p.A_adjoint_A(x), R.mri_AtA(x, mask)) < 1e-6
    assert rel_err(p.normal_step(x, aty, 0.8), x - 0.8 * (R.mri_AtA(x, mask) - aty)) < 1e-6
    assert rel_err(p.prox_l2(z, ref_y, 0.7), R.mri_prox_l2(z, ref_y, mask, 0.7)) < 1e-6
    assert rel_err(p.V_adjoint(x), R.im_to_kspace(x)) < 1e-6 and rel_err(p.V(x), R.kspace_to_im(x)) < 1e-6
    # the pipelined kernels against the general tile passes on the same operands
    got = p.A_adjoint(ref_y)
    monkeypatch.setenv("DINVK_NO_PIPE_FFT", "1")
    assert rel_err(got, p.A_adjoint(ref_y)) < 5e-7


def test_pipe320_single_and_multicoil(emul_backend):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    gen = torch.Generator().manual_seed(9)
    B, N, H, W = 2, 3, 320, 320
    x = torch.randn(B, 2, H, W, generator=gen)
    mask = _masks("lines", B, H, W, gen)
    p = dinv.physics.MRI(mask=mask, img_size=(2, H, W))
    y = R.mri_A(x, mask)
    assert rel_err(p.A(x), y) < 1e-6 and rel_err(p.A_adjoint(y), R.mri_At(y, mask)) < 1e-6
    maps = torch.randn(B, N, H, W, generator=gen, dtype=torch.complex64)
    maps = maps / maps.abs().pow(2).sum(1, keepdim=True).sqrt()
    pm = dinv.physics.MultiCoilMRI(mask=mask, coil_maps=maps, img_size=(2, H, W))
    ym = R.mcmri_A(x, mask, maps)
    assert rel_err(pm.A(x), ym) < 1e-6
    assert rel_err(pm.A_adjoint(ym), R.mcmri_At(ym, mask, maps)) < 1e-6
    assert rel_err(pm.A_adjoint(ym, rss=True), R.mcmri_At(ym, mask, maps, use_rss=True)) < 1e-6


@pytest.mark.parametrize("W,circle", [(64, False), (128, True)])
def test_tiled_radon_forward_and_transpose(W, circle, emul_backend, monkeypatch):
    """the shared-memory tiled Radon kernels (default GPU path for W >= 64): one tile (W = 64) and 2 x 2 tiles with the disc
    mask (W = 128) against the oracle and against the ray-per-thread / gather kernels on the same operands"""
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    gen = torch.Generator().manual_seed(11)
    angles = torch.tensor([0.0, 33.0, 90.0, 121.5, 170.0])
    x = torch.randn(1, 1, W, W, generator=gen)
    phys = dinv.physics.Tomography(angles=angles, img_width=W, circle=circle, normalize=False)
    y = phys.A(x)
    ref_y = R.tomography_A(x, angles, circle=circle)
    assert y.shape == ref_y.shape and rel_err(y, ref_y) < 1e-5
    v = torch.randn(ref_y.shape, generator=gen)
    xt = phys.A_adjoint(v)
    assert rel_err(xt, R.tomography_At(v, angles, W, circle=circle)) < 1e-5
    lhs, rhs = (y * v).sum().double(), (x * xt).sum().double()
    assert abs(float(lhs - rhs)) < 1e-4 * max(1.0, abs(float(lhs)))
    monkeypatch.setenv("DINVK_NO_TILED_RADON", "1")
    assert rel_err(phys.A(x), y) < 3e-6 and rel_err(phys.A_adjoint(v), xt) < 3e-6


def test_pipe256_every_operand_combination_vs_tile_passes(emul_backend, monkeypatch):
    """the operand / multiplier / epilogue combinations of `dinvk_spectral` that reach the pipelined kernels (the list of
    tests/test_gpu_spectral_pipe.py), pipelined kernels vs the general tile passes on the same operands, on the host"""
    import os

    from deepinv_b200 import _ffi, ops

    H = W = 256
    B = 2  # 32 tiles over the emulated persistent grid (8 CTAs): every CTA refills its ring
    gen = torch.Generator().manual_seed(2)
    r = lambda: torch.randn(B, 2, H, W, generator=gen)
    x, p1, q0, q1 = r(), r(), r(), r()
    line = ops.mask_spec_from_real((torch.rand(B, 1, 1, W, generator=gen) > 0.7).float().expand(B, 2, H, W).contiguous(), H, W)
    full = ops.mask_spec_from_real(torch.rand(B, 2, H, W, generator=gen), H, W)
    shared = ops.mask_spec_from_real(torch.rand(1, 2, H, W, generator=gen), H, W)
    assert line.sh == 0 and full.sh == W
    cb = torch.rand(B, generator=gen) + 0.5
    cases = []
    for mname, m in (("line", line), ("full", full), ("shared", shared)):
        for gmode in ((_ffi.G_MASK, _ffi.G_INV_SQ_PLUS_C, _ffi.G_PINV) if mname == "line" else (_ffi.G_MASK,)):
            cases.append((f"A {mname} g{gmode}", dict(fwd=True, inv=False, gmode=gmode, mask=m, c=0.8)))
            cases.append((f"At {mname} g{gmode}", dict(fwd=False, inv=True, gmode=gmode, mask=m, c=0.8)))
        cases.append((f"A {mname} epilogue", dict(fwd=True, inv=False, gmode=_ffi.G_MASK, mask=m, a0=0.5, p1=p1, a1=-1.5, e0=2.0,
                                                   q0=q0, e1=0.25, q1=q1, e2=-0.75)))
        cases.append((f"At {mname} c_batch", dict(fwd=False, inv=True, gmode=_ffi.G_INV_SQ_PLUS_C, mask=m, c_batch=cb)))
    for gmode in (_ffi.G_MASK, _ffi.G_SQ, _ffi.G_INV_SQ_PLUS_C, _ffi.G_PINV):
        cases.append((f"fused line g{gmode}", dict(fwd=True, inv=True, gmode=gmode, mask=line, c=1.3)))
    cases += [
        ("fused none", dict(fwd=True, inv=True)),
        ("A none", dict(fwd=True, inv=False)),
        ("At uncentred", dict(fwd=False, inv=True, centered=False, gmode=_ffi.G_MASK, mask=full)),
        ("fused uncentred", dict(fwd=True, inv=True, centered=False, gmode=_ffi.G_SQ, mask=line)),
        ("normal step", dict(fwd=True, inv=True, gmode=_ffi.G_SQ, mask=line, e0=-0.9, q0=x, e1=1.0, q1=q1, e2=0.9)),
        ("fused q0 != p0", dict(fwd=True, inv=True, gmode=_ffi.G_SQ, mask=line, e0=-0.9, q0=q0, e1=1.0, q1=q1, e2=0.9)),
        ("prox", dict(fwd=True, inv=True, gmode=_ffi.G_INV_SQ_PLUS_C, mask=line, p1=p1, a1=1.0 / 0.7, c=1.0 / 0.7)),
        ("prox c_batch", dict(fwd=True, inv=True, gmode=_ffi.G_INV_SQ_PLUS_C, mask=line, p1=p1, a1=1.0, c_batch=cb)),
    ]
    for name, kw in cases:
        got = ops.spectral(x, H, W, **kw)
        monkeypatch.setenv("DINVK_NO_PIPE_FFT", "1")
        want = ops.spectral(x, H, W, **kw)
        monkeypatch.delenv("DINVK_NO_PIPE_FFT")
        assert rel_err(got, want) < 3e-6, name

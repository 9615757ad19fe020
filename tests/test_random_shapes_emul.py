"""Randomised shape sweeps (hypothesis, fixed seed) of the SIMT kernels under host emulation against the oracle: sizes that are
not multiples of any tile, filters of every parity, per-sample filters, odd / prime transform lengths, channel counts that do not
fill a channel tile.  Index-math bugs live at such sizes; the fixtures and the benchmark sizes do not visit them."""
import pytest
import torch
from hypothesis import HealthCheck, given, seed, settings
from hypothesis import strategies as st

from conftest import rel_err

COMMON = dict(deadline=None, max_examples=20, suppress_health_check=list(HealthCheck), derandomize=True)


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


@settings(**COMMON)
@given(B=st.integers(1, 3), C=st.integers(1, 3), H=st.integers(5, 70), W=st.integers(5, 70), h=st.integers(1, 6), w=st.integers(1, 6),
       pad=st.sampled_from(["valid", "circular", "replicate", "reflect", "constant"]), per_sample=st.booleans(), per_channel=st.booleans())
def test_blur_forward_and_transpose(B, C, H, W, h, w, pad, per_sample, per_channel):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    h, w = min(h, H - 1), min(w, W - 1)
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + W)
    x = torch.randn(B, C, H, W, generator=g)
    filt = torch.rand(B if per_sample else 1, C if per_channel else 1, h, w, generator=g)
    phys = dinv.physics.Blur(filter=filt, padding=pad)
    y = R.blur_A(x, filt, pad)
    assert rel_err(phys.A(x), y) < 1e-5
    v = torch.randn(y.shape, generator=g)
    assert rel_err(phys.A_adjoint(v), R.blur_At(v, filt, pad, H, W)) < 1e-5


@settings(**COMMON)
@given(B=st.integers(1, 3), H=st.integers(1, 40), W=st.integers(1, 40), kind=st.sampled_from(["full", "lines", "shared"]),
       gamma=st.floats(0.2, 5.0))
def test_mri_all_transform_lengths(B, H, W, kind, gamma):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    g = torch.Generator().manual_seed(H * 41 + W)
    x = torch.randn(B, 2, H, W, generator=g)
    z = torch.randn(B, 2, H, W, generator=g)
    if kind == "lines":
        m = (torch.rand(B, 1, 1, W, generator=g) > 0.5).float().expand(B, 2, H, W).contiguous()
    elif kind == "shared":
        m = (torch.rand(1, 1, H, W, generator=g) > 0.5).float().expand(1, 2, H, W).contiguous()
    else:
        m = (torch.rand(B, 2, H, W, generator=g) > 0.5).float()
    p = dinv.physics.MRI(mask=m, img_size=(2, H, W))
    err = (lambda a, b: float((a - b).abs().max())) if H * W < 4 else rel_err
    y = R.mri_A(x, m)
    assert err(p.A(x), y) < 2e-6 and torch.equal(p.A(x) == 0, y == 0)
    assert err(p.A_adjoint(y), R.mri_At(y, m)) < 2e-6
    assert err(p.A_adjoint_A(x), R.mri_AtA(x, m)) < 2e-6
    assert err(p.prox_l2(z, y, gamma), R.mri_prox_l2(z, y, m, gamma)) < 1e-5
    aty = R.mri_At(y, m)
    assert err(p.normal_step(x, aty, 0.7), x - 0.7 * (R.mri_AtA(x, m) - aty)) < 2e-6


@settings(**COMMON)
@given(W=st.integers(4, 40), A=st.integers(1, 7), circle=st.booleans(), B=st.integers(1, 2))
def test_radon_forward_transpose_fbp(W, A, circle, B):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    g = torch.Generator().manual_seed(W * 7 + A)
    angles = torch.rand(A, generator=g) * 180
    x = torch.randn(B, 1, W, W, generator=g)
    phys = dinv.physics.Tomography(angles=angles, img_width=W, circle=circle, normalize=False)
    y = R.tomography_A(x, angles, circle=circle)
    assert rel_err(phys.A(x), y) < 1e-5
    v = torch.randn(y.shape, generator=g)
    assert rel_err(phys.A_adjoint(v), R.tomography_At(v, angles, W, circle=circle)) < 1e-5
    assert rel_err(phys.A_dagger(y, fbp=True), R.tomography_fbp(y, angles, W, circle=circle)) < 1e-5


@settings(**COMMON)
@given(B=st.integers(1, 2), cin=st.integers(1, 20), cout=st.integers(1, 40), H=st.integers(2, 20), W=st.integers(2, 45),
       kind=st.sampled_from([0, 1, 2]), relu=st.booleans())
def test_conv_forward_and_backward(B, cin, cout, H, W, kind, relu):
    import torch.nn.functional as F

    from deepinv_b200 import ops

    if kind == 1:
        H, W = 2 * ((H + 1) // 2), 2 * ((W + 1) // 2)
    g = torch.Generator().manual_seed(cin * 100 + cout)
    x = torch.randn(B, cin, H, W, generator=g, requires_grad=True)
    shape = (cin, cout, 2, 2) if kind == 2 else (cout, cin, 3, 3) if kind == 0 else (cout, cin, 2, 2)
    wt = (torch.randn(shape, generator=g) / 4).requires_grad_()
    bias = torch.randn(cout, generator=g, requires_grad=True)
    ref = {0: lambda t: F.conv2d(t, wt, bias, padding=1), 1: lambda t: F.conv2d(t, wt, bias, stride=2),
           2: lambda t: F.conv_transpose2d(t, wt, bias, stride=2)}[kind](x)
    ref = torch.relu(ref) if relu else ref
    out = ops.conv_f32_ag(x, wt, kind=kind, bias=bias, relu=relu)
    assert rel_err(out, ref) < 1e-5
    r = torch.randn(ref.shape, generator=g)
    got = torch.autograd.grad((out * r).sum(), [x, wt, bias])
    want = torch.autograd.grad((ref * r).sum(), [x, wt, bias])
    for a, b in zip(got, want):  # (a bias gradient is a sum with cancellation: measure it against the size of its terms)
        assert float((a - b).norm()) < 2e-5 * max(float(b.norm()), 1e-2 * float(r.norm()))


@settings(**COMMON)
@given(H=st.integers(4, 30), W=st.integers(4, 30), factor=st.sampled_from([1, 2, 3, 4]), filt=st.sampled_from([None, "gaussian", "bilinear", "bicubic", "sinc"]),
       pad=st.sampled_from(["circular", "reflect", "replicate", "constant"]))
def test_downsampling_shapes(H, W, factor, filt, pad):
    import deepinv_b200 as dinv
    from deepinv_b200.physics import functional as dF
    from oracle import ref_ops as R

    H, W = H - H % factor + factor, W - W % factor + factor
    g = torch.Generator().manual_seed(H * 31 + W)
    x = torch.randn(2, 2, H, W, generator=g)
    phys = dinv.physics.Downsampling(img_size=(2, H, W), filter=filt, factor=factor, padding=pad)
    f = {None: None, "gaussian": lambda: dF.gaussian_blur(sigma=(factor, factor)), "bilinear": lambda: dF.bilinear_filter(factor),
         "bicubic": lambda: dF.bicubic_filter(factor), "sinc": lambda: dF.sinc_filter(factor, length=4 * factor)}[filt]
    f = f() if f is not None else None
    if f is not None and pad in ("reflect",) and (f.shape[-1] // 2 >= W or f.shape[-2] // 2 >= H):
        return  # torch's reflect padding (and the reference) needs pad < size
    if f is not None and (f.shape[-1] > W or f.shape[-2] > H):
        return
    y = R.down_A(x, f, factor, pad)
    assert phys.A(x).shape == y.shape and rel_err(phys.A(x), y) < 1e-5
    v = torch.randn(y.shape, generator=g)
    assert rel_err(phys.A_adjoint(v), R.down_At(v, f, factor, pad, H, W)) < 1e-5


@settings(**COMMON)
@given(B=st.integers(1, 2), N=st.integers(1, 5), H=st.integers(2, 24), W=st.integers(2, 24), shared_maps=st.booleans(), rss=st.booleans())
def test_multicoil_shapes(B, N, H, W, shared_maps, rss):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    g = torch.Generator().manual_seed(N * 97 + H * 5 + W)
    x = torch.randn(B, 2, H, W, generator=g)
    maps = torch.randn(1 if shared_maps else B, N, H, W, generator=g, dtype=torch.complex64)
    maps = maps / maps.abs().pow(2).sum(1, keepdim=True).sqrt()
    m = (torch.rand(B, 1, 1, W, generator=g) > 0.4).float().expand(B, 2, H, W).contiguous()
    p = dinv.physics.MultiCoilMRI(mask=m, coil_maps=maps, img_size=(2, H, W))
    y = R.mcmri_A(x, m, maps)
    assert p.A(x).shape == y.shape and rel_err(p.A(x), y) < 2e-6
    assert rel_err(p.A_adjoint(y, rss=rss), R.mcmri_At(y, m, maps, use_rss=rss)) < 2e-6


@settings(**COMMON)
@given(B=st.integers(1, 3), C=st.integers(1, 3), H=st.integers(4, 33), W=st.integers(4, 33), h=st.integers(1, 5), w=st.integers(1, 5),
       gamma=st.floats(0.3, 4.0))
def test_blurfft_shapes(B, C, H, W, h, w, gamma):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    h, w = min(h, H), min(w, W)
    g = torch.Generator().manual_seed(H * 37 + W * 3 + h)
    filt = torch.rand(1, 1, h, w, generator=g) + 0.1
    filt = filt / filt.sum()
    x = torch.randn(B, C, H, W, generator=g)
    z = torch.randn(B, C, H, W, generator=g)
    p = dinv.physics.BlurFFT(img_size=(C, H, W), filter=filt)
    mask, angle = R.blurfft_params(filt, (C, H, W))
    y = R.blurfft_A(x, mask, angle, (C, H, W))
    assert rel_err(p.A(x), y) < 1e-5
    assert rel_err(p.A_adjoint(y), R.blurfft_At(y, mask, angle, (C, H, W))) < 1e-5
    assert rel_err(p.prox_l2(z, y, gamma), R.blurfft_prox_l2(z, y, mask, angle, (C, H, W), gamma)) < 1e-5


@settings(**COMMON)
@given(W=st.integers(6, 30), A=st.integers(1, 6), D=st.integers(3, 40), circle=st.booleans(), src=st.floats(20.0, 80.0), det=st.floats(10.0, 80.0),
       spacing=st.floats(0.03, 0.5))
def test_fanbeam_geometries(W, A, D, circle, src, det, spacing):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    g = torch.Generator().manual_seed(W * 13 + D)
    fp = {"n_detector_pixels": D, "detector_spacing": spacing, "source_radius": src, "detector_radius": det}
    angles = torch.rand(A, generator=g) * 360
    x = torch.randn(1, 1, W, W, generator=g)
    phys = dinv.physics.Tomography(angles=angles, img_width=W, circle=circle, fan_beam=True, fan_parameters=dict(fp), normalize=False)
    y = R.fanbeam_forward(x, angles, circle, dict(fp))
    got = phys.A(x)
    assert got.shape == y.shape
    # wide fans put most sample coordinates far outside [-1, 1]; the reference's own fp32 grid then carries errors of a few 1e-5
    # of the (small) output.  Yardstick: the fp64 evaluation — the kernel must be at least as close to it as the reference is
    y64 = R.fanbeam_forward(x.double(), angles.double(), circle, dict(fp))
    scale = max(float(y64.norm()), 1e-3 * float(x.norm()))
    assert float((got.double() - y64).norm()) < max(2e-5 * scale, 1.5 * float((y.double() - y64).norm()))
    v = torch.randn(y.shape, generator=g)
    xt = R.fanbeam_adjoint(v, angles, W, circle, dict(fp))
    xt64 = R.fanbeam_adjoint(v.double(), angles.double(), W, circle, dict(fp))
    scale = max(float(xt64.norm()), 1e-3 * float(v.norm()))
    assert float((phys.A_adjoint(v).double() - xt64).norm()) < max(2e-5 * scale, 1.5 * float((xt.double() - xt64).norm()))


@settings(**COMMON)
@given(B=st.integers(1, 2), T=st.integers(1, 4), H=st.integers(2, 18), W=st.integers(2, 18), form=st.sampled_from(["hw", "thw", "cthw", "bcthw"]))
def test_dynamic_mri_mask_forms(B, T, H, W, form):
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    g = torch.Generator().manual_seed(T * 50 + H * 3 + W)
    shape = {"hw": (H, W), "thw": (T, H, W), "cthw": (2, T, H, W), "bcthw": (B, 2, T, H, W)}[form]
    m_in = (torch.rand(*shape, generator=g) > 0.5).float()
    p = dinv.physics.DynamicMRI(mask=m_in, img_size=(2, T, H, W))
    m = m_in
    while m.dim() < 5:
        m = m.unsqueeze(0)
    if m.shape[1] == 1:
        m = torch.cat([m, m], 1)
    x = torch.randn(B, 2, p.mask.shape[2], H, W, generator=g)
    y = R.mri_A(x, m)
    assert rel_err(p.A(x), y) < 2e-6 and rel_err(p.A_adjoint(y), R.mri_At(y, m)) < 2e-6

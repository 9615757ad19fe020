"""Parity cases shared by the CUDA tests (`-m gpu`, real kernels) and the host-logic tests (CPU: the same
package code driving the emulated kernels).  Every case compares the package's public API with the golden
vectors produced by the real reference (tests/golden)."""
import math

import pytest

import torch

from conftest import load_golden, rel_err

TOL = 1e-5


def to_dev(g, dev):
    return {k: ({kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else v.to(dev)) for k, v in g.items()}


def load_model(cls, g, dev, **kw):
    m = cls(pretrained=None, device=dev, **kw)
    m.load_state_dict(g["sd"], strict=True)
    return m.eval()


def case_mri(name, dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    x, y, z, gam = g["x"], g["y"], g["z"], float(g["gamma"])
    phys = dinv.physics.MRI(mask=g["mask"], img_size=tuple(x.shape[1:]), device=dev)
    yk = phys.A(x)
    assert rel_err(yk, y) < TOL
    assert torch.equal(yk == 0, y == 0)
    assert rel_err(phys.A_adjoint(y), g["At"]) < TOL
    assert rel_err(phys.A_adjoint_A(x), g["AtA"]) < TOL
    assert rel_err(phys.A_A_adjoint(y), g["AAt"]) < TOL
    assert rel_err(phys.prox_l2(z, y, gam), g["prox"]) < TOL
    assert rel_err(phys.A_dagger(y), g["dagger"]) < TOL
    assert rel_err(phys.V_adjoint(x), g["Vt"]) < TOL
    assert rel_err(phys.V(x), g["V"]) < TOL
    assert rel_err(phys.A_adjoint(y, mag=True), g["At_mag"]) < TOL
    assert rel_err(phys.normal_step(x, phys.A_adjoint(y), 0.8), x - 0.8 * (g["AtA"] - g["At"])) < TOL
    phys2 = dinv.physics.MRI(img_size=tuple(x.shape[1:]), device=dev)  # mask= kwarg is stored (forward.py:249-276)
    assert rel_err(phys2.A(x, mask=g["mask"]), y) < TOL
    assert torch.equal(phys2.mask, g["mask"])
    xr = x.clone().requires_grad_(True)  # autograd: backward of A is A^T
    phys.A(xr).backward(y)
    assert rel_err(xr.grad, phys.A_adjoint(y)) < TOL


def case_dynamic_mri(name, dev):
    """DynamicMRI / SequentialMRI on the time-folded static kernels vs the real reference"""
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    x, y = g["x"], g["y"]
    if name.startswith("seq"):
        phys = dinv.physics.SequentialMRI(mask=g["mask_in"], img_size=tuple(g["mask"].shape[1:]), device=dev)
        assert torch.equal(phys.mask, g["mask"])
        yk = phys.A(x)
        assert yk.shape == y.shape and rel_err(yk, y) < TOL and torch.equal(yk == 0, y == 0)
        assert rel_err(phys.A_adjoint(y), g["At"]) < TOL
        assert rel_err(phys.A_adjoint(y, keep_time_dim=True), g["At_keep"]) < TOL
        assert rel_err(phys.A_dagger(y), g["dagger"]) < TOL
        xr = x.clone().requires_grad_(True)
        phys.A(xr).backward(y)
        assert rel_err(xr.grad, phys.A_adjoint(y, keep_time_dim=True).sum(2)) < TOL
        return
    phys = dinv.physics.DynamicMRI(mask=g["mask_in"], img_size=tuple(x.shape[1:]), device=dev)
    assert torch.equal(phys.mask, g["mask"])
    yk = phys.A(x)
    assert rel_err(yk, y) < TOL and torch.equal(yk == 0, y == 0)
    assert rel_err(phys.A_adjoint(y), g["At"]) < TOL
    assert rel_err(phys.A_adjoint(y, mag=True), g["At_mag"]) < TOL
    assert rel_err(phys.A_adjoint_A(x), g["AtA"]) < TOL
    assert rel_err(phys.prox_l2(g["z"], y, float(g["gamma"])), g["prox"]) < TOL
    assert rel_err(phys.A_dagger(y), g["dagger"]) < TOL
    assert rel_err(phys.normal_step(x, phys.A_adjoint(y), 0.8), x - 0.8 * (g["AtA"] - g["At"])) < TOL
    st = phys.to_static(device=dev)
    assert torch.equal(st.mask, torch.clip(g["mask"].sum(2), 0.0, 1.0))
    with pytest.raises(ValueError):
        phys.A(x[:, :, :2])


def case_filters(dev):
    from deepinv_b200.physics import functional as dF

    g = load_golden("filters")
    got = dict(gaussian=dF.gaussian_blur(sigma=(2.0, 2.0), device=dev), gaussian_aniso=dF.gaussian_blur(sigma=(1.0, 2.5), angle=30.0, device=dev),
               bilinear2=dF.bilinear_filter(2, device=dev), bicubic2=dF.bicubic_filter(2, device=dev),
               bicubic4=dF.bicubic_filter(4, device=dev), sinc2=dF.sinc_filter(2, length=8, device=dev),
               sinc3=dF.sinc_filter(3, length=12, device=dev))
    for k, v in got.items():
        assert v.shape == g[k].shape and rel_err(v, g[k]) < 1e-6, k


def case_downsampling(name, dev):
    """Downsampling: blur kernel + decimation, exact transpose, closed-form circular prox"""
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    _, filt, fac, pad = name.split("_")
    filt = None if filt == "none" else {"gauss": "gaussian"}.get(filt, filt)
    x = g["x"]
    phys = dinv.physics.Downsampling(img_size=tuple(x.shape[1:]), filter=filt, factor=int(g["factor"]), padding=pad, device=dev)
    y = phys.A(x)
    assert y.shape == g["y"].shape and rel_err(y, g["y"]) < TOL
    assert rel_err(phys.A_adjoint(g["v"]), g["At"]) < TOL
    u = torch.randn_like(x)
    lhs, rhs = (phys.A(u) * g["v"]).sum().double(), (u * phys.A_adjoint(g["v"])).sum().double()
    assert abs(float(lhs - rhs)) < 1e-4 * max(1.0, abs(float(lhs)))
    if "prox" in g:
        assert rel_err(phys.prox_l2(g["z"], g["y"], float(g["gamma"])), g["prox"]) < TOL
        cg = phys.prox_l2(g["z"], g["y"], float(g["gamma"]), use_fft=False)  # the CG route agrees with the closed form
        assert rel_err(cg, g["prox"]) < 1e-3
    xr = x.clone().requires_grad_(True)
    phys.A(xr).backward(g["v"])
    assert rel_err(xr.grad, g["At"]) < TOL


def case_combine(dev):
    """§8(b): composition (`*`) and stacking (`stack`) of operators — member kernels + CG prox vs the real reference"""
    import deepinv_b200 as dinv

    g = to_dev(load_golden("combine_down_blur"), dev)
    x = g["x"]
    blur = dinv.physics.Blur(filter=g["filt"], padding="circular", device=dev)
    down = dinv.physics.Downsampling(img_size=tuple(x.shape[1:]), filter="bilinear", factor=2, padding="circular", device=dev)
    comp = down * blur
    assert isinstance(comp, dinv.physics.ComposedLinearPhysics) and comp[0] is blur and comp[1] is down
    assert rel_err(comp.A(x), g["y"]) < TOL
    assert rel_err(comp.A_adjoint(g["v"]), g["At"]) < TOL
    assert rel_err(comp.prox_l2(g["z"], g["y"], float(g["gamma"])), g["prox"]) < 1e-4

    g = to_dev(load_golden("combine_stack_mri"), dev)
    x = g["x"]
    p1 = dinv.physics.MRI(mask=g["m1"], img_size=tuple(x.shape[1:]), device=dev)
    p2 = dinv.physics.MRI(mask=g["m2"], img_size=tuple(x.shape[1:]), device=dev)
    st = p1.stack(p2)
    assert isinstance(st, dinv.physics.StackedLinearPhysics) and len(st) == 2
    y = st.A(x)
    assert rel_err(y[0], g["y0"]) < TOL and rel_err(y[1], g["y1"]) < TOL
    assert rel_err(st.A_adjoint(y), g["At"]) < TOL
    assert rel_err(st.prox_l2(g["z"], y, float(g["gamma"])), g["prox"]) < 1e-4
    assert rel_err(st.A_dagger(y), g["dagger"]) < 5e-3
    yy = st(x)  # forward = sensor(noise(A)) per member
    assert isinstance(yy, dinv.physics.TensorList) and rel_err(yy[1], g["y1"]) < TOL


def case_mri_3d(dev):
    """single-coil 3-D MRI (three_d=True): separable 3-D transform on the 2-D kernels vs the real reference"""
    import deepinv_b200 as dinv

    g = to_dev(load_golden("mri3d_6x8x12"), dev)
    x, y = g["x"], g["y"]
    phys = dinv.physics.MRI(mask=g["mask_in"], img_size=tuple(x.shape[1:]), three_d=True, device=dev)
    assert torch.equal(phys.mask, g["mask"])
    yk = phys.A(x)
    assert rel_err(yk, y) < TOL and torch.equal(yk == 0, y == 0)
    assert rel_err(phys.V_adjoint(x), g["Vt"]) < TOL
    assert rel_err(phys.A_adjoint(y), g["At"]) < TOL
    assert rel_err(phys.A_adjoint_A(x), g["AtA"]) < TOL
    assert rel_err(phys.prox_l2(g["z"], y, float(g["gamma"])), g["prox"]) < TOL
    assert rel_err(phys.A_dagger(y), g["dagger"]) < TOL
    xr = x.clone().requires_grad_(True)
    phys.A(xr).backward(y)
    assert rel_err(xr.grad, g["At"]) < TOL


def case_multicoil(name, dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    maps = torch.complex(g["maps_re"], g["maps_im"])
    phys = dinv.physics.MultiCoilMRI(mask=g["mask"], coil_maps=maps, img_size=tuple(g["x"].shape[1:]), device=dev)
    assert rel_err(phys.A(g["x"]), g["y"]) < TOL
    assert rel_err(phys.A_adjoint(g["y"]), g["At"]) < TOL
    if "At_rss" in g:
        assert rel_err(phys.A_adjoint(g["y"], rss=True), g["At_rss"]) < TOL
    if "dagger" in g:
        assert rel_err(phys.A_dagger(g["y"]), g["dagger"]) < 1e-4


def case_tomography(name, dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    circle = "circle" in name
    W = g["x"].shape[-1]
    phys = dinv.physics.Tomography(angles=g["angles"], img_width=W, circle=circle, normalize=False, device=dev)
    y = phys.A(g["x"])
    assert y.shape == g["y"].shape and not y.is_contiguous()  # the reference returns the transposed view
    assert rel_err(y, g["y"]) < TOL
    assert rel_err(phys.A_adjoint(g["v"]), g["At"]) < TOL
    assert rel_err(phys.A_dagger(g["y"], fbp=True), g["fbp"]) < TOL
    physb = dinv.physics.Tomography(angles=g["angles"], img_width=W, circle=circle, normalize=False,
                                    adjoint_via_backprop=False, device=dev)
    assert rel_err(physb.A_adjoint(g["v"]), g["At_irad"]) < TOL
    assert rel_err(physb.A_dagger(g["y"], fbp=True), g["fbp_irad"]) < TOL
    u, v = torch.randn_like(g["x"]), torch.randn_like(g["y"])
    lhs, rhs = (phys.A(u) * v).sum().double(), (u * phys.A_adjoint(v)).sum().double()
    assert abs(float(lhs - rhs)) < 1e-4 * max(1.0, abs(float(lhs)))  # exact transpose (reference asserts 1e-3)


def case_fanbeam(name, dev):
    """fan-beam Tomography: forward, exact transpose, FBP vs the real reference"""
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    W = g["x"].shape[-1]
    fp = None if "default" in name else {"n_detector_pixels": 37, "detector_spacing": 0.31, "source_radius": 40.0,
                                         "detector_radius": 25.0}
    phys = dinv.physics.Tomography(angles=g["angles"], img_width=W, circle="circle" in name, fan_beam=True, fan_parameters=fp,
                                   normalize=False, device=dev)
    y = phys.A(g["x"])
    assert y.shape == g["y"].shape and rel_err(y, g["y"]) < TOL
    assert rel_err(phys.A_adjoint(g["v"]), g["At"]) < TOL
    assert rel_err(phys.A_dagger(g["y"], fbp=True), g["fbp"]) < TOL
    u = torch.randn_like(g["x"])
    lhs, rhs = (phys.A(u) * g["v"]).sum().double(), (u * phys.A_adjoint(g["v"])).sum().double()
    assert abs(float(lhs - rhs)) < 1e-4 * max(1.0, abs(float(lhs)))


def case_tomography_normalised(dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden("tomo_16_norm"), dev)
    phys = dinv.physics.Tomography(angles=g["angles"], img_width=16, normalize=True, device=dev)
    # the reference's norm depends on the global RNG state at construction: compare ours loosely, then adopt theirs
    assert abs(float(phys.operator_norm) / float(g["operator_norm"]) - 1) < 2e-2
    phys.operator_norm = g["operator_norm"].clone()
    assert rel_err(phys.A(g["x"]), g["y"]) < TOL
    assert rel_err(phys.A_adjoint(g["y"]), g["At"]) < TOL
    assert rel_err(phys.A_dagger(g["y"], fbp=True), g["fbp"]) < TOL
    # unregularised CG on a rank-deficient system (see tests/test_oracle_golden.py): 50 iterations amplify operator differences
    # of 1e-6 — the kernels' fp64 sample coordinates vs the reference's fp32 ones — by ~1e4; the reference's own pseudo-inverse
    # test allows 5 % (tests/test_physics.py:946-968)
    assert rel_err(phys.A_dagger(g["y"]), g["dagger"]) < 2e-2


def case_blur(name, dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    pad = name.split("_")[-1]
    phys = dinv.physics.Blur(filter=g["filt"], padding=pad, device=dev)
    assert rel_err(phys.A(g["x"]), g["y"]) < TOL
    assert rel_err(phys.A_adjoint(g["v"]), g["At"]) < TOL


def case_blur_cg(dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden("blur_gauss_circular_prox"), dev)
    phys = dinv.physics.Blur(filter=g["filt"], padding="circular", device=dev)
    assert rel_err(phys.prox_l2(g["z"], g["y"], float(g["gamma"])), g["prox"]) < 1e-4
    assert rel_err(phys.A_dagger(g["y"]), g["dagger"]) < 5e-3


def case_blurfft(name, dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden(name), dev)
    x = g["x"]
    phys = dinv.physics.BlurFFT(img_size=tuple(x.shape[1:]), filter=g["filt"], device=dev)
    assert rel_err(phys.mask, g["mask"]) < TOL
    assert rel_err(phys.A(x), g["y"]) < TOL
    assert rel_err(phys.A_adjoint(g["y"]), g["At"]) < TOL
    assert rel_err(phys.A_adjoint_A(x), g["AtA"]) < TOL
    assert rel_err(phys.prox_l2(g["z"], g["y"], float(g["gamma"])), g["prox"]) < TOL
    assert rel_err(phys.V_adjoint(x), g["Vt"]) < TOL
    assert rel_err(phys.V(phys.V_adjoint(x)), x) < TOL
    blur = dinv.physics.Blur(filter=g["filt"], padding="circular", device=dev)  # reference test_blur: Blur == BlurFFT
    assert rel_err(blur.A(x), g["y"]) < TOL
    # The remaining checks involve the PHASE of h^ and 1/|h^|.  Where |h^| ~ 1e-8 (sigma=2 has such bins) the phase of
    # an fp32 FFT is round-off noise and 1/|h^| (hard threshold 1e-5) amplifies 1e-7 differences by 1e5, so two correct
    # FFTs of the filter disagree there.  These checks therefore run on the reference's own spectral buffers, set
    # through the public buffers exactly like `update_parameters(mask=...)` would.
    phys.mask = g["mask"].clone()
    phys.angle = torch.complex(g["angle_re"], g["angle_im"])
    assert rel_err(phys.U_adjoint(x), g["Ut"]) < TOL
    assert rel_err(phys.U(phys.mask * phys.V_adjoint(x)), g["y"]) < TOL
    # pseudo-inverse: round-off of y^ itself (1e-7) is amplified by up to 1e5 in the smallest retained bins; the
    # reference's own pseudo-inverse test uses 5 % (tests/test_physics.py:946-968); better-conditioned fixtures: 1e-4
    assert rel_err(phys.A_dagger(g["y"]), g["dagger"]) < (5e-2 if "cfg1" in name else 1e-4)


def case_drunet(dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden("drunet_tiny"), dev)
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2)
    with torch.no_grad():
        assert rel_err(den(g["x"], 0.05), g["out"]) < TOL
        assert rel_err(den(g["x"], g["sig"]), g["out_b"]) < TOL
        assert rel_err(den(g["xs"], 0.05), g["out_s"]) < TOL


def case_dncnn(dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden("dncnn_tiny"), dev)
    den = load_model(dinv.models.DnCNN, g, dev, in_channels=1, out_channels=1, depth=5, nf=8)
    with torch.no_grad():
        assert rel_err(den(g["x"], 0.1), g["out"]) < TOL


def case_optim_toy(dev):
    """Step algebra of every iterator (PGD + relaxation, FISTA, HQS, ADMM, DRS f-/g-first, GD + RED) and of the DDRM sampler
    with a closed-form 'denoiser': the package loops on the kernels vs the oracle loops.  (The same algorithms with the real
    DRUNet against the reference's golden outputs run on the GPU; the host emulation runs them with the real network only
    for PGD — a denoiser pass costs seconds there.)"""
    import deepinv_b200 as dinv
    from oracle import ref_ops as R

    from deepinv_b200.optim import ADMM, DRS, FISTA, GD, HQS, L2, PGD, RED, PnP

    g = to_dev(load_golden("optim_mri_tiny"), dev)
    m, y = g["mask"], g["y"]
    phys = dinv.physics.MRI(mask=m, img_size=(2, 32, 32), device=dev)
    toy = lambda v, s: v * (1.0 / (1.0 + float(s)))
    mc, yc = m.cpu(), y.cpu()
    A, At = (lambda v: R.mri_A(v, mc)), (lambda v: R.mri_At(v, mc))
    prox = lambda v, gam: R.mri_prox_l2(v, yc, mc, gam)
    kw = dict(data_fidelity=L2(), prior=PnP(toy), early_stop=False)
    relax = PGD(max_iter=3, params_algo={"stepsize": 0.8, "g_param": 0.05, "lambda": 1.0, "beta": 0.9}, **kw)
    assert rel_err(relax(y, phys), R.pgd(yc, A, At, toy, 0.8, 0.05, 3, beta=0.9)) < TOL
    assert rel_err(FISTA(stepsize=1.0, sigma_denoiser=0.05, max_iter=4, **kw)(y, phys), R.fista(yc, A, At, toy, 1.0, 0.05, 4)) < TOL
    assert rel_err(HQS(stepsize=0.9, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys), R.hqs(yc, prox, At, toy, 0.9, 0.05, 3)) < TOL
    assert rel_err(ADMM(stepsize=1.1, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys), R.admm(yc, prox, At, toy, 1.1, 0.05, 3)) < TOL
    assert rel_err(DRS(stepsize=1.0, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys), R.drs(yc, prox, At, toy, 1.0, 0.05, 3)) < TOL
    gfirst = DRS(max_iter=3, g_first=True, params_algo={"stepsize": 0.7, "g_param": 0.05, "lambda": 1.0, "beta": 0.8}, **kw)
    assert rel_err(gfirst(y, phys), R.drs(yc, prox, At, toy, 0.7, 0.05, 3, beta=0.8, g_first=True)) < TOL
    red = GD(data_fidelity=L2(), prior=RED(toy), stepsize=0.5, lambda_reg=0.3, sigma_denoiser=0.05, max_iter=3, early_stop=False)
    assert rel_err(red(y, phys), R.gd(yc, A, At, lambda v: v - toy(v, 0.05), 0.5, 0.3, 3)) < TOL
    gd = to_dev(load_golden("ddrm_mri_tiny"), dev)
    physn = dinv.physics.MRI(mask=gd["mask"], img_size=(2, 32, 32), device=dev,
                             noise_model=dinv.physics.GaussianNoise(sigma=float(gd["sigma_noise"])))
    sig = [float(s_) for s_ in gd["sigmas"]]
    out = dinv.sampling.DDRM(denoiser=toy, sigmas=sig)(gd["y"], physn, noises=list(gd["noises"]))
    want = R.ddrm(gd["y"].cpu(), lambda v: v, R.kspace_to_im, R.im_to_kspace, gd["mask"].cpu(), toy, sig, list(gd["noises"].cpu()),
                  sigma_noise=float(gd["sigma_noise"]))
    assert rel_err(out, want) < TOL


def case_pnp_mri(dev, full=True):
    import deepinv_b200 as dinv
    from deepinv_b200.optim import ADMM, FISTA, HQS, L2, PGD, PnP

    g = to_dev(load_golden("optim_mri_tiny"), dev)
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2)
    phys = dinv.physics.MRI(mask=g["mask"], img_size=(2, 32, 32), device=dev)
    y = g["y"]
    kw = dict(data_fidelity=L2(), prior=PnP(den), early_stop=False)
    assert rel_err(PGD(stepsize=1.0, sigma_denoiser=0.05, max_iter=4, **kw)(y, phys), g["pgd"]) < TOL
    if full:  # (the slow host emulation runs PGD with the real network + case_optim_toy; every algorithm runs on the GPU)
        assert rel_err(ADMM(stepsize=1.0, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys), g["admm"]) < TOL
        relax = PGD(max_iter=3, params_algo={"stepsize": 0.8, "g_param": 0.05, "lambda": 1.0, "beta": 0.9}, **kw)
        assert rel_err(relax(y, phys), g["pgd_relax"]) < TOL
        assert rel_err(HQS(stepsize=1.0, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys), g["hqs"]) < TOL
        assert rel_err(FISTA(stepsize=1.0, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys), g["fista"]) < TOL
    x, m = PGD(stepsize=1.0, sigma_denoiser=0.05, max_iter=2, **kw)(y, phys, compute_metrics=True, x_gt=g["x"])
    assert len(m["residual"]) == y.shape[0] and len(m["residual"][0]) == 2 and len(m["psnr"][0]) == 3


def case_drs_gd_dpir(dev, full=True):
    """DRS / GD / DPIR on the kernels of the PGD path (full=False: the subset the slow host
    emulation runs; every case runs on the GPU)"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import DPIR, DRS, GD, L2, RED, PnP, Tikhonov

    g = to_dev(load_golden("optim2_mri_tiny"), dev)
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2)
    phys = dinv.physics.MRI(mask=g["mask"], img_size=(2, 32, 32), device=dev)
    y = g["y"]
    kw = dict(data_fidelity=L2(), early_stop=False)
    assert rel_err(GD(prior=Tikhonov(), stepsize=0.5, lambda_reg=0.1, max_iter=4, **kw)(y, phys), g["gd_tik"]) < TOL
    if not full:
        return
    assert rel_err(DRS(prior=PnP(den), stepsize=1.0, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys), g["drs"]) < TOL
    relax = DRS(prior=PnP(den), max_iter=3, g_first=True,
                params_algo={"stepsize": 0.7, "g_param": 0.05, "lambda": 1.0, "beta": 0.8}, **kw)
    assert rel_err(GD(prior=RED(den), stepsize=0.5, lambda_reg=0.3, sigma_denoiser=0.05, max_iter=3, **kw)(y, phys),
                   g["gd_red"]) < TOL
    assert rel_err(relax(y, phys), g["drs_relax"]) < TOL
    # DPIR's first proxes use gamma up to 64, where (A^T y + z/gamma)/(s^2 + 1/gamma) loses ~3e-6 per prox in fp32 for any
    # implementation (the reference's own fp32 result is 1e-5 from the fp64 evaluation, tests/test_host_logic_emul.py)
    assert rel_err(DPIR(sigma=0.05, denoiser=den, device=dev)(y, phys), g["dpir"]) < 5e-5
    physb = dinv.physics.Blur(filter=g["filt"], padding="circular", device=dev)
    assert rel_err(DPIR(sigma=0.05, denoiser=den, device=dev)(g["yb"], physb), g["dpir_blur"]) < 1e-4  # CG prox inside


def _check_param_grads(model, g, tol_w=2e-5, tol_p=2e-4):
    """gradients of every parameter against the reference's (fixture keys `grad__<name>`); the reference registers the
    unfolded parameters under `init_params_algo` (BaseUnfold) or `params_algo` (BaseOptim): same tensors"""
    want = {k[6:].replace("__", ".").replace("init_params_algo", "params_algo"): v for k, v in g.items() if k.startswith("grad__")}
    scale = max(float(v.abs().max()) for k, v in want.items() if k.startswith("params_algo"))
    seen = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        assert k in want, k
        seen += 1
        if k.startswith("params_algo"):
            assert abs(float(p.grad) - float(want[k])) < tol_p * scale, (k, float(p.grad), float(want[k]))
        else:
            assert rel_err(p.grad, want[k]) < tol_w, (k, rel_err(p.grad, want[k]))
    assert seen == len(want), (seen, len(want))


def case_train_deq_explicit(dev):
    """deep equilibrium GD + Tikhonov on MRI: forward value, loss and d loss / d (stepsize, lambda) through the backward
    fixed-point hook == the real reference"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import GD, L2, DEQConfig, Tikhonov

    g0 = to_dev(load_golden("train_unfolded_pgd_mri"), dev)
    g = to_dev(load_golden("train_deq_gd_tikhonov"), dev)
    phys = dinv.physics.MRI(mask=g0["mask"], img_size=(2, 32, 32), device=dev)
    deq = GD(data_fidelity=L2(), prior=Tikhonov(), stepsize=0.5, lambda_reg=0.2, max_iter=10, early_stop=False,
             DEQ=DEQConfig(max_iter_backward=12), trainable_params=["stepsize", "lambda"]).to(dev)
    out = deq(g0["y"], phys)
    loss = ((out - g0["x"]) ** 2).mean()
    loss.backward()
    assert rel_err(out, g["out"]) < TOL and abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    _check_param_grads(deq, g)


def case_train_unfolded(dev):
    """one training step (forward, MSE loss, backward) of unfolded PGD (MRI, DRUNet), a deep-equilibrium PGD and unfolded
    ADMM (circular Blur: CG prox with implicit-differentiation backward, DnCNN): outputs, losses and the gradient of
    every trainable parameter == the real reference"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import PGD, L2, DEQConfig, PnP
    from deepinv_b200.unfolded import unfolded_builder

    g = to_dev(load_golden("train_unfolded_pgd_mri"), dev)
    phys = dinv.physics.MRI(mask=g["mask"], img_size=(2, 32, 32), device=dev)
    x, y = g["x"], g["y"]
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2).train()
    model = unfolded_builder("PGD", params_algo={"stepsize": [1.0, 0.8], "g_param": [0.05, 0.03], "lambda": 1.0},
                             trainable_params=["stepsize", "g_param"], data_fidelity=L2(), prior=PnP(den), max_iter=2).to(dev)
    out = model(y, phys)
    loss = ((out - x) ** 2).mean()
    loss.backward()
    assert rel_err(out, g["out"]) < TOL and abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    _check_param_grads(model, g)

    gd = to_dev(load_golden("train_deq_pgd_mri"), dev)
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2).train()
    deq = PGD(data_fidelity=L2(), prior=PnP(den), stepsize=0.9, sigma_denoiser=0.05, max_iter=6, early_stop=False,
              DEQ=DEQConfig(max_iter_backward=8), trainable_params=["stepsize"]).to(dev)
    out = deq(y, phys)
    loss = ((out - x) ** 2).mean()
    loss.backward()
    assert rel_err(out, gd["out"]) < TOL and abs(float(loss) - float(gd["loss"])) < 1e-5 * float(gd["loss"])
    _check_param_grads(deq, gd, tol_w=5e-5)

    gb = to_dev(load_golden("train_unfolded_admm_blur"), dev)
    physb = dinv.physics.Blur(filter=gb["filt"], padding="circular", device=dev)
    dn = load_model(dinv.models.DnCNN, gb, dev, in_channels=1, out_channels=1, depth=5, nf=8).train()
    modelb = unfolded_builder("ADMM", params_algo={"stepsize": [1.0, 1.2], "g_param": 0.05, "lambda": 1.0, "beta": 1.0},
                              trainable_params=["stepsize"], data_fidelity=L2(), prior=PnP(dn), max_iter=2).to(dev)
    out = modelb(gb["y"], physb)
    loss = ((out - gb["x"]) ** 2).mean()
    loss.backward()
    assert rel_err(out, gb["out"]) < 1e-4 and abs(float(loss) - float(gb["loss"])) < 1e-4 * float(gb["loss"])
    _check_param_grads(modelb, gb, tol_w=2e-3, tol_p=2e-3)  # CG solves (tol 1e-4) inside forward and backward


def case_anderson(dev, full=True):
    """Anderson-accelerated loops (fixed_point.py:117-260) vs the real reference"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import GD, L2, PGD, AndersonAccelerationConfig, PnP, Tikhonov

    g = to_dev(load_golden("optim_anderson"), dev)
    phys = dinv.physics.MRI(mask=g["mask"], img_size=(2, 32, 32), device=dev)
    gd = GD(data_fidelity=L2(), prior=Tikhonov(), stepsize=0.5, lambda_reg=0.2, max_iter=12, early_stop=False,
            anderson_acceleration=AndersonAccelerationConfig(history_size=3, beta=1.0, eps=1e-3))
    assert rel_err(gd(g["y"], phys), g["gd"]) < 2e-5  # 12 small linear solves on top of the iterates
    from deepinv_b200.optim import BacktrackingConfig

    # starts too large: the first step is rejected and the stepsize halved.  8 iterations only: every accept / reject decision
    # then has a > 2x margin; run to convergence, the decrease F(x_prev) - F(x) drops to the fp32 round-off of F and the
    # decisions (hence the final stepsize) become noise on ANY implementation
    bt = GD(data_fidelity=L2(), prior=Tikhonov(), stepsize=2.5, lambda_reg=0.2, max_iter=8, early_stop=False,
            backtracking=BacktrackingConfig(gamma=0.1, eta=0.5, max_iter=20))
    assert rel_err(bt(g["y"], phys), g["gd_bt"]) < TOL
    assert abs(float(bt.init_params_algo["stepsize"][0]) - float(g["gd_bt_step"])) < 1e-6
    pbt = PGD(data_fidelity=L2(), prior=Tikhonov(), stepsize=3.0, lambda_reg=0.5, max_iter=10, early_stop=False, backtracking=True)
    assert rel_err(pbt(g["y"], phys), g["pgd_bt"]) < TOL
    assert abs(float(pbt.init_params_algo["stepsize"][0]) - float(g["pgd_bt_step"])) < 1e-6
    if not full:
        return
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2)
    pgd = PGD(data_fidelity=L2(), prior=PnP(den), stepsize=1.0, sigma_denoiser=0.05, max_iter=4, early_stop=False,
              anderson_acceleration=True)
    assert rel_err(pgd(g["y"], phys), g["pgd"]) < 2e-5


def case_pdcp(dev):
    """Chambolle-Pock primal-dual iterations (primal_dual_CP.py): f-first with K = identity, g-first with K = A, vs the reference"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import L2, PDCP, Tikhonov

    g = to_dev(load_golden("optim_anderson"), dev)
    gp = to_dev(load_golden("optim_pdcp"), dev)
    phys = dinv.physics.MRI(mask=g["mask"], img_size=(2, 32, 32), device=dev)
    kwp = dict(data_fidelity=L2(), prior=Tikhonov(), lambda_reg=0.3, stepsize=0.6, stepsize_dual=0.8, max_iter=6, early_stop=False)
    assert rel_err(PDCP(**kwp)(g["y"], phys), gp["cp"]) < TOL
    assert rel_err(PDCP(g_first=True, K=phys.A, K_adjoint=phys.A_adjoint, **kwp)(g["y"], phys), gp["cp_gfirst"]) < TOL


def case_pnp_blur_admm(dev):
    import deepinv_b200 as dinv
    from deepinv_b200.optim import ADMM, L2, PnP

    g = to_dev(load_golden("optim_blur_tiny"), dev)
    den = load_model(dinv.models.DnCNN, g, dev, in_channels=1, out_channels=1, depth=5, nf=8)
    phys = dinv.physics.Blur(filter=g["filt"], padding="circular", device=dev)
    out = ADMM(data_fidelity=L2(), prior=PnP(den), stepsize=1.0, sigma_denoiser=0.05, max_iter=3, early_stop=False)(g["y"], phys)
    assert rel_err(out, g["admm"]) < 1e-4  # three CG solves inside


def case_ddrm_inpainting(dev):
    """DDRM on Inpainting and Denoising (identity singular vectors): the reference's own sampling demo, recorded noise draws"""
    import deepinv_b200 as dinv

    g = to_dev(load_golden("ddrm_inpainting_tiny"), dev)
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2)
    sig = float(g["sigma_noise"])
    phys = dinv.physics.Inpainting(img_size=(2, 32, 32), mask=g["mask"], device=dev, noise_model=dinv.physics.GaussianNoise(sigma=sig))
    assert rel_err(phys.A_adjoint(g["y"]), g["At"]) < TOL and rel_err(phys.prox_l2(g["x"], g["y"], 0.8), g["prox"]) < TOL
    sigmas = [float(s_) for s_ in g["sigmas"]]
    out = dinv.sampling.DDRM(denoiser=den, sigmas=sigmas)(g["y"], phys, noises=list(g["noises"]))
    assert rel_err(out, g["out"]) < TOL
    dn = dinv.physics.Denoising(dinv.physics.GaussianNoise(sigma=sig), device=dev)
    out = dinv.sampling.DDRM(denoiser=den, sigmas=sigmas)(g["y"], dn, noises=list(g["noises"]))
    assert rel_err(out, g["out_denoising"]) < TOL


def case_diffpir(dev):
    """DiffPIR on BlurFFT with the reference's recorded noise draws (host-planned schedule vs the reference's device lookups)"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import L2

    g = to_dev(load_golden("diffpir_blurfft_tiny"), dev)
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=1, out_channels=1, nc=(8, 16, 32, 64), nb=2)
    phys = dinv.physics.BlurFFT(img_size=(1, 32, 32), filter=g["filt"], device=dev,
                                noise_model=dinv.physics.GaussianNoise(sigma=float(g["sigma_noise"])))
    model = dinv.sampling.DiffPIR(den, L2(), sigma=0.03, max_iter=6, zeta=0.3, lambda_=7.0, device=dev)
    out = model(g["y"], phys, noises=list(g["noises"]))
    # clamp(-1, 1) of the denoised estimate is a non-smooth step: entries within round-off of the bounds may land on either
    # side; everything else is the usual 1e-5
    assert rel_err(out, g["out"]) < 5e-5


def case_ddrm(dev):
    import deepinv_b200 as dinv

    g = to_dev(load_golden("ddrm_mri_tiny"), dev)
    den = load_model(dinv.models.DRUNet, g, dev, in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2)
    phys = dinv.physics.MRI(mask=g["mask"], img_size=(2, 32, 32), device=dev,
                            noise_model=dinv.physics.GaussianNoise(sigma=float(g["sigma_noise"])))
    model = dinv.sampling.DDRM(denoiser=den, sigmas=[float(s) for s in g["sigmas"]])
    out = model(g["y"], phys, noises=list(g["noises"]))
    assert rel_err(out, g["out"]) < TOL

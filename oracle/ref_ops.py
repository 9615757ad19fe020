"""oracle/ref_ops.py — CPU restatement of the reference algorithms on the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under deepinv_b200/ imports this module; only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs do.  It restates, in
plain torch-CPU tensor code (dtype follows the inputs: float32 like the reference, or float64 as a
tighter yardstick), what deepinv v0.4.1 computes for each row of SURVEY.md §8(a).  Every function
cites the reference lines it follows.

Pinning: tests/test_oracle_golden.py checks every function here against vectors produced by the
REAL reference imported in the authoring container (tests/golden/make_golden.py; fixtures committed
under tests/golden/*.npz), so parity claims against this oracle are anchored on the reference.

The reference's arithmetic lives in PyTorch ATen (pocketfft/MKL FFT, mkldnn conv, grid_sampler);
torch>=2.2 is its pinned dependency (pyproject.toml:29, here 2.11.0).  FFTs and dense convolutions
below call the same ATen CPU routines the reference calls (that IS the reference CPU path, and it is
what the bench's CPU baseline must time); Radon sampling, padding/folding, CG, step algebra and DDRM
are restated explicitly.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# -------------------------------------------------------------------------------------------------
# a1/a2: planar <-> complex and the centred orthonormal DFT  (deepinv/utils/mixins.py:148-206)
# -------------------------------------------------------------------------------------------------


def to_complex(x: torch.Tensor) -> torch.Tensor:
    """(B,2,...,H,W) real -> (B,...,H,W) complex  (mixins.py:148-151)"""
    return torch.view_as_complex(x.movedim(1, -1).contiguous())


def from_complex(x: torch.Tensor) -> torch.Tensor:
    """(B,...,H,W) complex -> (B,2,...,H,W) real  (mixins.py:153-156)"""
    return torch.view_as_real(x).movedim(-1, 1)


def cfft2(xc: torch.Tensor) -> torch.Tensor:
    """fftshift(fftn(ifftshift(x), ortho))  (mixins.py:170-180)"""
    return torch.fft.fftshift(torch.fft.fftn(torch.fft.ifftshift(xc, dim=(-2, -1)), dim=(-2, -1), norm="ortho"), dim=(-2, -1))


def cifft2(xc: torch.Tensor) -> torch.Tensor:
    """fftshift(ifftn(ifftshift(x), ortho))  (mixins.py:158-168)"""
    return torch.fft.fftshift(torch.fft.ifftn(torch.fft.ifftshift(xc, dim=(-2, -1)), dim=(-2, -1), norm="ortho"), dim=(-2, -1))


def im_to_kspace(x):  # mixins.py:182-193
    return from_complex(cfft2(to_complex(x)))


def kspace_to_im(y):  # mixins.py:195-206
    return from_complex(cifft2(to_complex(y)))


def cfft3(xc: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """three_d=True: the same centred orthonormal transform over the last THREE dims (mixins.py:158-180, dim=(-3,-2,-1))"""
    d = (-3, -2, -1)
    f = torch.fft.ifftn if inverse else torch.fft.fftn
    return torch.fft.fftshift(f(torch.fft.ifftshift(xc, dim=d), dim=d, norm="ortho"), dim=d)


def im_to_kspace3(x):
    return from_complex(cfft3(to_complex(x)))


def kspace_to_im3(y):
    return from_complex(cfft3(to_complex(y), inverse=True))


def check_mask(mask: torch.Tensor) -> torch.Tensor:
    """to (B,2,H,W), duplicating a real mask on both planes (mixins.py:125-146)"""
    while mask.dim() < 4:
        mask = mask.unsqueeze(0)
    if mask.shape[1] == 1:
        mask = torch.cat([mask, mask], dim=1)
    return mask


# -------------------------------------------------------------------------------------------------
# a3-a5: MRI as a DecomposablePhysics  (physics/forward.py:1080-1252, physics/mri.py:100-135)
# -------------------------------------------------------------------------------------------------


def mri_A(x, mask):  # forward.py:1095 with U = id, V^T = im_to_kspace
    return mask * im_to_kspace(x)


def mri_At(y, mask):  # forward.py:1116
    return kspace_to_im(torch.conj(mask) * y)


def mri_AtA(x, mask):  # forward.py:1140
    return kspace_to_im(mask.conj() * mask * im_to_kspace(x))


def mri_AAt(y, mask):  # forward.py:1128
    return mask.conj() * mask * y


def mri_prox_l2(z, y, mask, gamma):  # forward.py:1212-1234
    b = mri_At(y, mask) + 1 / gamma * z
    scaling = torch.conj(mask) * mask + 1 / gamma
    return kspace_to_im(im_to_kspace(b) / scaling)


def mri_dagger(y, mask):  # forward.py:1236-1252
    m = torch.where(mask > 1e-5, mask.reciprocal(), torch.zeros_like(mask))
    return kspace_to_im(y * m)


def rss(x, multicoil=True, mag=True):  # mixins.py:249-287
    ss = x.pow(2)
    if mag:
        ss = ss.sum(dim=1, keepdim=True)
    if multicoil:
        ss = ss.sum(dim=2)
    return ss.sqrt()


# DynamicMRI / SequentialMRI (physics/mri.py:499-695): the functions above act on the last two dims, so a (B,2,T,H,W) video
# with a (B|1,2,T,H,W) mask goes through mri_A / mri_At / mri_AtA / mri_prox_l2 unchanged (the reference folds T into B)
def seqmri_A(x, mask5):  # mri.py:672-676: repeat the static image over T, then the dynamic operator
    return mri_A(x.unsqueeze(2).expand(x.shape[0], 2, *mask5.shape[2:]), mask5)


def time_average(x, mask=None):  # mixins.py:83-101
    _x = x.sum(2)
    m = (mask if mask is not None else (x != 0)).sum(2)
    out = torch.zeros_like(_x)
    out[m != 0] = _x[m != 0] / m[m != 0]
    return out


def seqmri_At(y, mask5):  # mri.py:678-695 (keep_time_dim=False): static adjoint of the time-averaged k-space
    return mri_At(time_average(y, mask5), time_average(mask5))


# a6: MultiCoilMRI (physics/mri.py:254-324)
def mcmri_A(x, mask, coil_maps):
    Sx = coil_maps * to_complex(x)[:, None]  # (B,N,H,W)
    return mask[:, :, None] * from_complex(cfft2(Sx))  # (B,2,N,H,W)


def mcmri_At(y, mask, coil_maps, use_rss=False):
    My = to_complex(mask[:, :, None] * y)
    FiMy = cifft2(My)
    if use_rss:
        return rss(from_complex(FiMy), multicoil=True)
    return from_complex(torch.sum(torch.conj(coil_maps) * FiMy, dim=1))


# -------------------------------------------------------------------------------------------------
# a7-a9: Radon / IRadon / ramp filter  (physics/functional/radon.py, physics/tomography.py)
# -------------------------------------------------------------------------------------------------


def deg2rad(theta: torch.Tensor) -> torch.Tensor:
    """x * 4 * atan(1) / 180 in the tensor's dtype (radon.py:70-71)"""
    return theta * 4 * torch.ones(1, dtype=theta.dtype).atan() / 180


def default_angles(n: int) -> torch.Tensor:
    return torch.linspace(0, 180, steps=n + 1)[:-1]  # tomography.py:136-139


def padded_width(W: int) -> int:
    """P = ceil(sqrt(2) * W) evaluated in float32 (radon.py:60-61, 262-263, 319)"""
    sqrt2 = (2 * torch.ones(1)).sqrt()
    return int((sqrt2 * W).ceil())


def pad_before(W: int, P: int) -> int:
    return (W + (P - W)) // 2 - W // 2  # radon.py:262-267


def _bilinear_zeros(img: torch.Tensor, px: torch.Tensor, py: torch.Tensor) -> torch.Tensor:
    """bilinear sampling of img (BC,P,P) at pixel coords (px,py) (same for every image), zeros outside —
    the arithmetic of grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True)."""
    P_h, P_w = img.shape[-2:]
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    wx1 = px - x0
    wy1 = py - y0
    wx0 = 1 - wx1
    wy0 = 1 - wy1
    x0 = x0.long()
    y0 = y0.long()
    out = torch.zeros(img.shape[0], *px.shape, dtype=img.dtype)
    flat = img.reshape(img.shape[0], -1)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi < P_w) & (yi >= 0) & (yi < P_h)
            idx = (yi.clamp(0, P_h - 1) * P_w + xi.clamp(0, P_w - 1)).reshape(-1)
            v = flat[:, idx].reshape(img.shape[0], *px.shape)
            out = out + v * (wx * wy * ok.to(img.dtype))
    return out


def radon_geometry(W: int, circle: bool):
    P = W if circle else padded_width(W)
    pb = 0 if circle else pad_before(W, P)
    return P, pb


def radon_forward(x: torch.Tensor, angles_deg: torch.Tensor, circle: bool = False) -> torch.Tensor:
    """(B,C,W,W) -> (B,C,P,A): rotate-sample-accumulate of Radon.forward (radon.py:252-309) with the
    affine_grid(align_corners=True) sampling grid of _create_grids (:311-342)."""
    B, C, W, _ = x.shape
    P, pb = radon_geometry(W, circle)
    dt = x.dtype
    if circle:
        ax = 2 * torch.arange(W, dtype=torch.float32) / (W - 1) - 1.0  # radon.py:271-281
        disc = ((ax[None, :] ** 2 + ax[:, None] ** 2) <= 1).to(dt)
        xp = x * disc
    else:
        xp = F.pad(x, (pb, P - W - pb, pb, P - W - pb))
    img = xp.reshape(B * C, P, P)
    lin = torch.linspace(-1, 1, P, dtype=dt)
    xj = lin[None, :].expand(P, P)  # base grid x (varies along width j)
    yi = lin[:, None].expand(P, P)  # base grid y (varies along height i)
    out = torch.zeros(B * C, P, len(angles_deg), dtype=dt)
    for t, th in enumerate(angles_deg):
        th = deg2rad(th.to(dt).reshape(1))
        c, s = th.cos(), th.sin()
        gx = c * xj + s * yi
        gy = -s * xj + c * yi
        px = (gx + 1) / 2 * (P - 1)
        py = (gy + 1) / 2 * (P - 1)
        out[:, :, t] = _bilinear_zeros(img, px, py).sum(1)  # sum over rows i -> (BC, P_j)
    return out.reshape(B, C, P, len(angles_deg))


def fan_parameters_default(W: int, fan_parameters=None) -> dict:
    """radon.py:224-240"""
    fp = dict(fan_parameters or {})
    fp.setdefault("pixel_spacing", 0.5 / W)
    fp.setdefault("source_radius", 57.5)
    fp.setdefault("detector_radius", 57.5)
    fp.setdefault("n_detector_pixels", 258)
    fp.setdefault("detector_spacing", 0.077)
    return fp


def fanbeam_forward(x: torch.Tensor, angles_deg: torch.Tensor, circle: bool = False, fan_parameters=None) -> torch.Tensor:
    """(B,C,W,W) -> (B,C,D,A): Radon.forward with fan_beam=True — the sampling grid of fan_beam_grid (radon.py:16-52): the
    points (x_i, y_j * d_i), d_i = 0.5 L (x_i + r_s) / (r_s + r_d) in units scaled by 2 / (G * pixel_spacing), rotated by theta;
    bilinear samples (align_corners=True, zeros) summed over i.  Differentiable: its vjp is the reference's adjoint."""
    B, C, W, _ = x.shape
    G, pb = radon_geometry(W, circle)
    fp = fan_parameters_default(W, fan_parameters)
    dt = x.dtype
    if circle:
        ax = 2 * torch.arange(W, dtype=torch.float32) / (W - 1) - 1.0
        xp = x * ((ax[None, :] ** 2 + ax[:, None] ** 2) <= 1).to(dt)
    else:
        xp = F.pad(x, (pb, G - W - pb, pb, G - W - pb))
    img = xp.reshape(B * C, G, G)
    D = int(fp["n_detector_pixels"])
    sf = 2.0 / (G * fp["pixel_spacing"])
    rs, rd, sp = fp["source_radius"] * sf, fp["detector_radius"] * sf, fp["detector_spacing"] * sf
    L = sp * (D - 1)
    xi = torch.linspace(-1, 1, G, dtype=dt)[None, :].expand(D, G)   # along the ray (grid width)
    yj = torch.linspace(-1, 1, D, dtype=dt)[:, None].expand(D, G)   # detector coordinate (grid height)
    yy = yj * (0.5 * L * (xi + rs) / (rs + rd))
    outs = []
    for th in angles_deg:
        th = deg2rad(th.to(dt).reshape(1))
        c, s = th.cos(), th.sin()
        px = ((c * xi + s * yy) + 1) / 2 * (G - 1)
        py = ((-s * xi + c * yy) + 1) / 2 * (G - 1)
        outs.append(_bilinear_zeros(img, px, py).sum(2))  # sum over the ray samples i -> (BC, D)
    return torch.stack(outs, dim=-1).reshape(B, C, D, len(angles_deg))


def fanbeam_adjoint(y: torch.Tensor, angles_deg: torch.Tensor, W: int, circle: bool = False, fan_parameters=None) -> torch.Tensor:
    """the autograd transpose the reference uses for fan-beam (tomography.py:322-342)"""
    x0 = torch.zeros(y.shape[0], y.shape[1], W, W, dtype=y.dtype, requires_grad=True)
    with torch.enable_grad():
        out = fanbeam_forward(x0, angles_deg, circle, fan_parameters)
    return torch.autograd.grad(out, x0, y)[0]


def radon_adjoint(y: torch.Tensor, angles_deg: torch.Tensor, W: int, circle: bool = False) -> torch.Tensor:
    """exact transpose of radon_forward (the autograd adjoint of Tomography, tomography.py:322-342):
    every sample scatters sino[b,c,j,theta]*w into its bilinear neighbours, then crop (pad^T)."""
    B, C, P, A = y.shape
    dt = y.dtype
    _, pb = radon_geometry(W, circle)
    lin = torch.linspace(-1, 1, P, dtype=dt)
    xj = lin[None, :].expand(P, P)
    yi = lin[:, None].expand(P, P)
    acc = torch.zeros(B * C, P * P, dtype=dt)
    yy = y.reshape(B * C, P, A)
    for t, th in enumerate(angles_deg):
        th = deg2rad(th.to(dt).reshape(1))
        c, s = th.cos(), th.sin()
        px = ((c * xj + s * yi) + 1) / 2 * (P - 1)
        py = ((-s * xj + c * yi) + 1) / 2 * (P - 1)
        x0, y0 = torch.floor(px), torch.floor(py)
        wx1, wy1 = px - x0, py - y0
        x0, y0 = x0.long(), y0.long()
        val = yy[:, :, t][:, None, :].expand(B * C, P, P)  # value of detector j broadcast over rows i
        for dy, wy in ((0, 1 - wy1), (1, wy1)):
            for dx, wx in ((0, 1 - wx1), (1, wx1)):
                xi, yi_ = x0 + dx, y0 + dy
                ok = (xi >= 0) & (xi < P) & (yi_ >= 0) & (yi_ < P)
                idx = (yi_.clamp(0, P - 1) * P + xi.clamp(0, P - 1)).reshape(-1)
                w = (wx * wy * ok.to(dt)).reshape(1, -1)
                acc.index_add_(1, idx, val.reshape(B * C, -1) * w)
    full = acc.reshape(B, C, P, P)
    if circle:
        ax = 2 * torch.arange(W, dtype=torch.float32) / (W - 1) - 1.0
        disc = ((ax[None, :] ** 2 + ax[:, None] ** 2) <= 1).to(dt)
        return full * disc
    return full[:, :, pb: pb + W, pb: pb + W]


def ramp_filter(y: torch.Tensor) -> torch.Tensor:
    """AbstractFilter.forward + RampFilter along dim -2 of (B,C,N,A) (radon.py:79-173)"""
    N = y.shape[-2]
    L = max(64, int(2 ** (2 * torch.tensor(N)).float().log2().ceil()))
    # (the index vector takes the dtype of y: in the fp32 evaluation nothing changes, and an fp64 evaluation — the yardstick of
    # the full-size parity tests — gets the kernel coefficients in fp64 instead of fp32-rounded ones)
    n = torch.cat([torch.arange(1, L / 2 + 1, 2), torch.arange(L / 2 - 1, 0, -2)]).to(y.dtype)
    f = torch.zeros(L, dtype=y.dtype)
    f[0] = 0.25
    f[1::2] = -1 / (torch.pi * n) ** 2
    ff = (2 * torch.fft.rfft(f, dim=-1)).unsqueeze(-1)
    padded = F.pad(y, (0, 0, 0, L - N))
    return torch.fft.irfft(torch.fft.rfft(padded, dim=-2) * ff, dim=-2)[:, :, :N, :].contiguous()


def iradon_backproject(y: torch.Tensor, angles_deg: torch.Tensor, W: int, circle: bool = False) -> torch.Tensor:
    """IRadon.forward(filtering=False) (radon.py:396-450): sinogram sampled bilinearly at
    (X = 2t/(A-1)-1, T = x cos - y sin), summed over angles, cropped, * pi/(2A)."""
    B, C, P, A = y.shape
    dt = y.dtype
    lin = torch.linspace(-1, 1, P, dtype=dt)
    ygrid, xgrid = torch.meshgrid(lin, lin, indexing="ij")
    sino = y.reshape(B * C, P, A)  # "image" of height P (detector) and width A (angle)
    reco = torch.zeros(B * C, P, P, dtype=dt)
    for t, th in enumerate(angles_deg):
        th = deg2rad(th.to(dt).reshape(1))
        T = xgrid * th.cos() - ygrid * th.sin()
        X = torch.ones(P, P, dtype=dt) * t * 2.0 / (A - 1) - 1.0
        px = (X + 1) / 2 * (A - 1)
        py = (T + 1) / 2 * (P - 1)
        reco = reco + _bilinear_zeros(sino, px, py)
    reco = reco.reshape(B, C, P, P)
    if not circle:
        pb = pad_before(W, P)
        reco = reco[:, :, pb: pb + W, pb: pb + W]
    else:
        disc = (xgrid ** 2 + ygrid ** 2) <= 1
        reco = reco * disc.to(dt)
    return reco * torch.pi / (2 * A)


def tomography_A(x, angles_deg, circle=False, operator_norm=None):  # tomography.py:238-256
    out = radon_forward(x, angles_deg, circle)
    return out / operator_norm if operator_norm is not None else out


def tomography_At(y, angles_deg, W, circle=False, operator_norm=None, via_backprop=True):  # tomography.py:311-350
    if via_backprop:
        out = radon_adjoint(y, angles_deg, W, circle)
    else:
        out = iradon_backproject(y, angles_deg, W, circle) / torch.pi * (2 * len(angles_deg))  # radon.py:512-514
    return out / operator_norm if operator_norm is not None else out


def tomography_fbp(y, angles_deg, W, circle=False, operator_norm=None, via_backprop=True):  # tomography.py:258-293
    A = len(angles_deg)
    yf = ramp_filter(y)
    if via_backprop:
        out = tomography_At(yf, angles_deg, W, circle, operator_norm, True) * torch.pi / (2 * A)
        if operator_norm is not None:
            out = out * operator_norm ** 2
    else:
        out = iradon_backproject(yf, angles_deg, W, circle)
        if operator_norm is not None:
            out = out * operator_norm
    return out


# -------------------------------------------------------------------------------------------------
# a10: Blur direct convolution and its transpose  (physics/functional/convolution.py:42-164, 641-758)
# -------------------------------------------------------------------------------------------------
_PAD_MODES = ("valid", "circular", "replicate", "reflect", "constant")


def _ext_index(a: torch.Tensor, n: int, mode: str):
    """map an extended index onto [0,n) under the padding rule; returns (index, valid_mask)"""
    if mode == "circular":
        return a % n, torch.ones_like(a, dtype=torch.bool)
    if mode == "replicate":
        return a.clamp(0, n - 1), torch.ones_like(a, dtype=torch.bool)
    if mode == "reflect":
        r = torch.where(a < 0, -a, a)
        r = torch.where(r > n - 1, 2 * (n - 1) - r, r)
        return r, torch.ones_like(a, dtype=torch.bool)
    ok = (a >= 0) & (a < n)
    return a.clamp(0, n - 1), ok


def _expand_filter(filt, B, C):
    b, c = filt.shape[:2]
    assert c in (1, C), f"Number of channels of the kernel is not matched for broadcasting, got c={c} and C={C}"
    assert b in (1, B), f"Batch size of the kernel is not matched for broadcasting, got b={b} and B={B}"
    return filt.expand(B, C, *filt.shape[2:])


def blur_A(x: torch.Tensor, filt: torch.Tensor, padding: str = "valid") -> torch.Tensor:
    """true convolution out[i,j] = sum k[u,v] x~[i-u+h//2, j-v+w//2] (conv2d, convolution.py:42-107;
    valid: out[i,j] = sum k[u,v] x[i+h-1-u, j+w-1-v])."""
    if padding == "zeros":
        padding = "constant"
    if padding not in _PAD_MODES:
        raise ValueError(f"padding = '{padding}' not implemented.")
    B, C, H, W = x.shape
    k = _expand_filter(filt, B, C)
    h, w = k.shape[-2:]
    if padding == "valid":
        Ho, Wo = H - h + 1, W - w + 1
        out = torch.zeros(B, C, Ho, Wo, dtype=x.dtype)
        for u in range(h):
            for v in range(w):
                out = out + k[:, :, u, v][:, :, None, None] * x[:, :, h - 1 - u: h - 1 - u + Ho, w - 1 - v: w - 1 - v + Wo]
        return out
    ph, pw = h // 2, w // 2
    ii = torch.arange(H)
    jj = torch.arange(W)
    out = torch.zeros_like(x)
    for u in range(h):
        ri, rok = _ext_index(ii - u + ph, H, padding)
        for v in range(w):
            ci, cok = _ext_index(jj - v + pw, W, padding)
            patch = x[:, :, ri][:, :, :, ci] * (rok[:, None] & cok[None, :]).to(x.dtype)
            out = out + k[:, :, u, v][:, :, None, None] * patch
    return out


def blur_At(y: torch.Tensor, filt: torch.Tensor, padding: str, H: int, W: int) -> torch.Tensor:
    """exact transpose of blur_A (conv_transpose2d + _apply_transpose_padding, convolution.py:110-164,
    689-758), written as the literal scatter-transpose of the gather above."""
    if padding == "zeros":
        padding = "constant"
    B, C = y.shape[:2]
    k = _expand_filter(filt, B, C)
    h, w = k.shape[-2:]
    out = torch.zeros(B, C, H, W, dtype=y.dtype)
    if padding == "valid":
        Ho, Wo = H - h + 1, W - w + 1
        for u in range(h):
            for v in range(w):
                out[:, :, h - 1 - u: h - 1 - u + Ho, w - 1 - v: w - 1 - v + Wo] += k[:, :, u, v][:, :, None, None] * y
        return out
    ph, pw = h // 2, w // 2
    ii = torch.arange(H)
    jj = torch.arange(W)
    flat = out.reshape(B, C, H * W)
    for u in range(h):
        ri, rok = _ext_index(ii - u + ph, H, padding)
        for v in range(w):
            ci, cok = _ext_index(jj - v + pw, W, padding)
            idx = (ri[:, None] * W + ci[None, :]).reshape(-1)
            contrib = (k[:, :, u, v][:, :, None, None] * y * (rok[:, None] & cok[None, :]).to(y.dtype)).reshape(B, C, -1)
            flat.index_add_(2, idx, contrib)
    return flat.reshape(B, C, H, W)


# Downsampling (physics/blur.py:280-364): blur, then keep every factor-th pixel; transpose = zero-stuffing + transposed blur
def down_A(x, filt, factor, padding="circular"):
    xb = x if filt is None else blur_A(x, filt, padding)
    return xb[:, :, ::factor, ::factor]


def down_At(y, filt, factor, padding, H, W):
    Hb, Wb = (H - filt.shape[-2] + 1, W - filt.shape[-1] + 1) if (filt is not None and padding == "valid") else (H, W)
    v = torch.zeros(y.shape[0], y.shape[1], Hb, Wb, dtype=y.dtype)
    v[:, :, ::factor, ::factor] = y
    return v if filt is None else blur_At(v, filt, padding, H, W)


def down_prox_l2(z, y, filt, factor, gamma):
    """closed form for circular padding (blur.py:332-364, Zhao et al. 2016)"""
    B, C, H, W = z.shape
    Fh = filter_fft(filt.expand(filt.shape[0], C, *filt.shape[-2:]) if filt.shape[1] != C else filt, (C, H, W), real_fft=False)
    z_hat = down_At(y, filt, factor, "circular", H, W) + z / gamma
    Fz = torch.fft.fft2(z_hat)
    fold = lambda a: a.reshape(*a.shape[:-2], factor, H // factor, factor, W // factor).mean(dim=(-4, -2))
    top, below = fold(Fh * Fz), fold(Fh.conj() * Fh) + 1 / gamma
    r = torch.real(torch.fft.ifft2(Fh.conj() * (top / below).repeat(1, 1, factor, factor)))
    return (z_hat - r) * gamma


# a11: BlurFFT (physics/blur.py:639-692, convolution.py:790-812)
def filter_fft(filt: torch.Tensor, img_size, real_fft: bool = True) -> torch.Tensor:
    H, W = img_size[-2:]
    h, w = filt.shape[-2:]
    f = F.pad(filt, (0, W - w, 0, H - h))
    f = torch.roll(f, shifts=(-int(h / 2), -int(w / 2)), dims=(-2, -1))
    return torch.fft.rfftn(f, dim=(-2, -1)) if real_fft else torch.fft.fftn(f, dim=(-2, -1))


def blurfft_params(filt: torch.Tensor, img_size):
    """mask = |h^| duplicated on a trailing axis, angle = exp(i arg h^)  (blur.py:659-692)"""
    if img_size[0] > filt.shape[1]:
        filt = filt.repeat(1, img_size[0], 1, 1)
    hf = filter_fft(filt, img_size)
    angle = torch.exp(1.0j * torch.angle(hf))
    m = torch.abs(hf).unsqueeze(-1)
    return torch.cat([m, m], dim=-1), angle


def blurfft_Vt(x):
    return torch.view_as_real(torch.fft.rfft2(x, norm="ortho"))


def blurfft_V(xb, img_size):
    return torch.fft.irfft2(torch.view_as_complex(xb.contiguous()), norm="ortho", s=tuple(img_size[-2:]))


def blurfft_U(xb, angle, img_size):
    return torch.fft.irfft2(torch.view_as_complex(xb.contiguous()) * angle, norm="ortho", s=tuple(img_size[-2:]))


def blurfft_Ut(x, angle):
    return torch.view_as_real(torch.fft.rfft2(x, norm="ortho") * torch.conj(angle))


def blurfft_A(x, mask, angle, img_size):
    return blurfft_U(mask * blurfft_Vt(x), angle, img_size)


def blurfft_At(y, mask, angle, img_size):
    return blurfft_V(torch.conj(mask) * blurfft_Ut(y, angle), img_size)


def blurfft_prox_l2(z, y, mask, angle, img_size, gamma):
    b = blurfft_At(y, mask, angle, img_size) + 1 / gamma * z
    return blurfft_V(blurfft_Vt(b) / (mask.conj() * mask + 1 / gamma), img_size)


def blurfft_dagger(y, mask, angle, img_size):
    m = torch.where(mask > 1e-5, mask.reciprocal(), torch.zeros_like(mask))
    return blurfft_V(blurfft_Ut(y, angle) * m, img_size)


# -------------------------------------------------------------------------------------------------
# a12: CG on the normal equations  (optim/linear/conjugate_gradient.py:35-77, least_squares.py:148-151)
# -------------------------------------------------------------------------------------------------


def _bdot(a, b):
    return (a.conj() * b).reshape(a.shape[0], -1).sum(-1).reshape((-1,) + (1,) * (a.dim() - 1))


def conjugate_gradient(Aop, b, max_iter=100, tol=1e-5, eps=1e-8, init=None):
    x = torch.zeros_like(b) if init is None else init
    r = b - Aop(x)
    p = r
    res_old = _bdot(r, r).real
    b_norm_sq = _bdot(b, b).real
    b_norm_sq = torch.where(b_norm_sq > 0, b_norm_sq, torch.ones_like(b_norm_sq))
    tolv = b_norm_sq * (tol ** 2)
    n_it = 0
    for i in range(int(max_iter)):
        Ap = Aop(p)
        alpha = res_old / (_bdot(p, Ap) + eps)
        x = x + p * alpha
        r = r - Ap * alpha
        res_new = _bdot(r, r).real
        n_it = i + 1
        if torch.all(res_new < tolv):
            break
        p = r + p * (res_new / (res_old + eps))
        res_old = res_new
        if i > 0 and i % 100 == 0:
            r = b - Aop(x)
            res_old = _bdot(r, r).real
    return x, n_it


def prox_l2_cg(A, At, z, y, gamma, max_iter=50, tol=1e-4, init=None):
    """LinearPhysics.prox_l2 with solver='CG' (forward.py:751-814 -> least_squares.py:148-151)"""
    b = At(y) + z / gamma
    H = lambda v: At(A(v)) + v / gamma
    return conjugate_gradient(H, b, max_iter=max_iter, tol=tol, init=z if init is None else init)


# -------------------------------------------------------------------------------------------------
# a16: denoisers  (models/drunet.py:200-263, 400-433; models/dncnn.py:116-131)
# -------------------------------------------------------------------------------------------------


def _resblock(x, sd, prefix):
    r = F.conv2d(x, sd[prefix + ".res.0.weight"], padding=1)
    r = F.conv2d(F.relu(r), sd[prefix + ".res.2.weight"], padding=1)
    return x + r


def drunet_forward_unet(x0, sd, nb=4):
    x1 = F.conv2d(x0, sd["m_head.weight"], padding=1)

    def down(x, name):
        for i in range(nb):
            x = _resblock(x, sd, f"{name}.{i}")
        return F.conv2d(x, sd[f"{name}.{nb}.weight"], stride=2)

    def up(x, name):
        x = F.conv_transpose2d(x, sd[f"{name}.0.weight"], stride=2)
        for i in range(1, nb + 1):
            x = _resblock(x, sd, f"{name}.{i}")
        return x

    x2 = down(x1, "m_down1")
    x3 = down(x2, "m_down2")
    x4 = down(x3, "m_down3")
    x = x4
    for i in range(nb):
        x = _resblock(x, sd, f"m_body.{i}")
    x = up(x + x4, "m_up3")
    x = up(x + x3, "m_up2")
    x = up(x + x2, "m_up1")
    return F.conv2d(x + x1, sd["m_tail.weight"], padding=1)


def drunet_forward(x, sigma, sd, nb=4):
    """noise-level channel + U-Net; sizes that are not multiples of 8 or < 32 are replicate-padded to a
    multiple of 16 (drunet.py:212-263, models/utils.py:49-61)"""
    B, _, H, W = x.shape
    if isinstance(sigma, torch.Tensor) and sigma.dim() > 0:
        nl = sigma.reshape(B, 1, 1, 1).expand(-1, 1, H, W).to(x.dtype) if sigma.numel() == B else sigma
    else:
        nl = torch.full((B, 1, H, W), float(sigma), dtype=x.dtype)
    x0 = torch.cat((x, nl), 1)
    if H % 8 == 0 and W % 8 == 0 and H > 31 and W > 31:
        return drunet_forward_unet(x0, sd, nb)
    ph, pw = int(math.ceil(H / 16) * 16 - H), int(math.ceil(W / 16) * 16 - W)
    return drunet_forward_unet(F.pad(x0, (0, pw, 0, ph), mode="replicate"), sd, nb)[..., :H, :W]


def dncnn_forward(x, sd, depth=20):
    t = F.relu(F.conv2d(x, sd["in_conv.weight"], sd.get("in_conv.bias"), padding=1))
    for i in range(depth - 2):
        t = F.relu(F.conv2d(t, sd[f"conv_list.{i}.weight"], sd.get(f"conv_list.{i}.bias"), padding=1))
    return F.conv2d(t, sd["out_conv.weight"], sd.get("out_conv.bias"), padding=1) + x


# -------------------------------------------------------------------------------------------------
# a13-a15: step algebra and loop drivers  (optim/optim_iterators/*.py, optimizers.py:572, fixed_point.py:324-359)
# -------------------------------------------------------------------------------------------------


def pgd(y, A, At, denoiser, stepsize, sigma_d, max_iter, lam=1.0, sigma_f=1.0, beta=1.0):
    """PGD with L2 + PnP, g_first=False (pgd.py:137-168, optim_iterator.py:112-132)"""
    x = At(y)
    norm = 1.0 / sigma_f ** 2
    for _ in range(max_iter):
        grad = norm * (At(A(x)) - At(y))  # data_fidelity.py:335-336 (A^T y recomputed as in the reference)
        z = x - stepsize * grad
        xn = denoiser(z, sigma_d)
        x = beta * xn + (1 - beta) * x if beta != 1.0 else xn
    return x


def fista(y, A, At, denoiser, stepsize, sigma_d, max_iter, a=3):
    x = At(y)
    z = At(y)
    for k in range(max_iter):
        alpha = (k + a - 1) / (k + a)
        zz = z - stepsize * (At(A(z)) - At(y))
        xn = denoiser(zz, sigma_d)
        z = xn + alpha * (xn - x)
        x = xn
    return x


def admm(y, prox_f, At, denoiser, stepsize, sigma_d, max_iter, beta=1.0):
    """ADMM, g_first=False (admm.py:58-68,108-147): u = prox_f(x - z), x = D(u + z), z += beta (u - x)"""
    x = At(y)
    z = At(y)
    for _ in range(max_iter):
        u = prox_f(x - z, stepsize)
        x = denoiser(u + z, sigma_d)
        z = z + beta * (u - x)
    return x


def hqs(y, prox_f, At, denoiser, stepsize, sigma_d, max_iter):
    x = At(y)
    for _ in range(max_iter):
        x = denoiser(prox_f(x, stepsize), sigma_d)
    return x


def drs(y, prox_f, At, denoiser, stepsize, sigma_d, max_iter, beta=1.0, g_first=False):
    """Douglas-Rachford splitting (optim_iterators/drs.py:36-73): x0 = z0 = A^T y"""
    x = z = At(y)
    for _ in range(max_iter):
        if g_first:
            u = denoiser(z, sigma_d)
            x = prox_f(2 * u - z, stepsize)
        else:
            u = prox_f(z, stepsize)
            x = denoiser(2 * u - z, sigma_d)
        z = z + beta * (x - u)
    return x


def gd(y, A, At, grad_g, stepsize, lam, max_iter):
    """gradient descent (optim_iterators/gradient_descent.py:48-56): x <- x - gamma (lambda grad g(x) + A^T(Ax - y))"""
    x = At(y)
    for _ in range(max_iter):
        x = x - stepsize * (lam * grad_g(x) + (At(A(x)) - At(y)))
    return x


def dpir_params(noise_level_img):
    """optim/dpir.py:11-35"""
    max_iter = 8
    s1, s2 = 49.0 / 255.0, noise_level_img
    sig = torch.logspace(torch.log10(torch.tensor(s1, dtype=torch.float32)), torch.log10(torch.tensor(s2, dtype=torch.float32)),
                         steps=max_iter, dtype=torch.float32)
    step = (sig / max(0.01, noise_level_img)) ** 2
    return sig, (1 / 0.23) * step, max_iter


def dpir(y, prox_f, At, denoiser, noise_level_img):
    """DPIR (optim/dpir.py:38-81): HQS with the per-iteration schedule above"""
    sig, step, n = dpir_params(noise_level_img)
    x = At(y)
    for k in range(n):
        x = denoiser(prox_f(x, step[k]), sig[k])
    return x


# -------------------------------------------------------------------------------------------------
# a17: DDRM  (sampling/diffusion.py:149-224) with the noise draws supplied by the caller
# -------------------------------------------------------------------------------------------------


def ddrm(y, Ut, V, Vt, mask1, denoiser, sigmas, noises, sigma_noise=0.01, eta=0.85, etab=1.0, eps=1e-6):
    """`noises[t]` replaces the t-th torch.randn_like draw; mask1 is the batch-1 mask (|.| taken here)"""
    B = y.shape[0]
    mask = torch.cat([mask1.abs()] * B, dim=0)
    c = math.sqrt(1 - eta ** 2)
    y_bar = Ut(y).clone()
    case = mask > sigma_noise
    y_bar[case] = y_bar[case] / (mask[case] + eps)
    nsr = torch.zeros_like(mask)
    nsr[case] = sigma_noise / (mask[case] + eps)
    mean = torch.zeros_like(y_bar)
    std = torch.ones_like(y_bar) * sigmas[0]
    mean[case] = y_bar[case]
    std[case] = (sigmas[0] ** 2 - nsr[case].pow(2)).sqrt()
    x_bar = mean + std * noises[0] / math.sqrt(2.0)
    x_bar_prev = x_bar
    x = denoiser(V(x_bar), sigmas[0])
    for t in range(1, len(sigmas)):
        x_bar = Vt(x)
        case2 = torch.logical_and(case, (sigmas[t] < nsr))
        case3 = torch.logical_and(case, (sigmas[t] >= nsr))
        mean = x_bar + c * sigmas[t] * (x_bar_prev - x_bar) / sigmas[t - 1]
        mean[case2] = x_bar[case2] + c * sigmas[t] * (y_bar[case2] - x_bar[case2]) / (nsr[case2] + eps)
        mean[case3] = (1.0 - etab) * x_bar[case3] + etab * y_bar[case3]
        std = torch.ones_like(x_bar) * eta * sigmas[t]
        std[case3] = (sigmas[t] ** 2 - (nsr[case3] * etab).pow(2)).clamp(min=0).sqrt()
        x_bar = mean + std * noises[t] / math.sqrt(2.0)
        x_bar_prev = x_bar
        x = denoiser(V(x_bar), sigmas[t])
    return x

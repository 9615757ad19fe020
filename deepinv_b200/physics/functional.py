"""Filter constructors used by Blur / BlurFFT / Downsampling (deepinv/physics/functional/blur.py:137-372,552-583).

Host-side parameter generation (a few hundred taps): plain torch on whatever device is asked for; the operators that
consume these filters run on the libdinvk kernels."""
from __future__ import annotations

from math import pi, sqrt

import torch


def gaussian_blur(psf_size=None, sigma=(1.0, 1.0), angle=0.0, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """2-D anisotropic, rotated Gaussian kernels (B,1,h,w), each normalised to sum 1 (functional/blur.py:137-264).
    sigma: float | (sy, sx) | (B,2) tensor in (height, width) order; angle: degrees, float | (B,) tensor."""
    if isinstance(sigma, (int, float)):
        sigma = (float(sigma), float(sigma))
    sig = torch.as_tensor(sigma, device=device, dtype=dtype)
    if sig.dim() == 1:
        sig = sig[None]
    if sig.shape[-1] != 2:
        raise ValueError("deepinv_b200.gaussian_blur: 2-D kernels only (sigma = (sy, sx))")
    ang = torch.as_tensor(angle, device=device, dtype=dtype).reshape(-1)
    B = max(sig.shape[0], ang.numel())
    sig, ang = sig.expand(B, 2), ang.expand(B)
    if psf_size is None:
        c = int(float(sig.max()) / 0.3 + 1)
        psf_size = (2 * c + 1, 2 * c + 1)
    h, w = psf_size
    ay = torch.linspace(-((h - 1) / 2), (h - 1) / 2, h, device=device, dtype=dtype)
    ax = torch.linspace(-((w - 1) / 2), (w - 1) / 2, w, device=device, dtype=dtype)
    yy, xx = torch.meshgrid(ay, ax, indexing="ij")
    coords = torch.stack([xx, yy], dim=-1)[None].expand(B, h, w, 2)
    th = ang * (pi / 180.0)
    rot = torch.stack([torch.cos(th), -torch.sin(th), torch.sin(th), torch.cos(th)], dim=1).view(B, 2, 2)
    coords = torch.einsum("bij,b...j->b...i", rot, coords)
    sxy = torch.flip(sig, dims=(1,))  # (sx, sy)
    kernel = torch.ones((B, h, w), device=device, dtype=dtype)
    for d in range(2):
        s = sxy[:, d].view(B, 1, 1)
        kernel = kernel * torch.exp(-0.5 * coords[..., d] ** 2 / s ** 2) / (sqrt(2 * pi) * s)
    kernel = kernel / kernel.sum(dim=(1, 2), keepdim=True)
    return kernel[:, None]


def kaiser_window(beta: float, length: int, device="cpu") -> torch.Tensor:
    if beta < 0:
        raise ValueError("beta must be greater than 0")
    if length < 1:
        raise ValueError("length must be greater than 0")
    if length == 1:
        return torch.tensor([1.0])
    half = (length - 1) / 2
    n = torch.arange(length, device=device)
    beta = torch.tensor(beta, device=device)
    return torch.i0(beta * torch.sqrt(1 - ((n - half) / half) ** 2)) / torch.i0(beta)


def sinc_filter(factor=2, length: int = 11, windowed: bool = True, device="cpu") -> torch.Tensor:
    """separable anti-aliasing sinc, optionally Kaiser-windowed (functional/blur.py:283-336)"""
    if isinstance(factor, torch.Tensor):
        factor = factor.cpu().item()
    deltaf = 2 * (2 - 1.4142136) / factor
    n = torch.arange(length, device=device) - (length - 1) / 2
    f = torch.sinc(n / factor)
    if windowed:
        A = 2.285 * (length - 1) * 3.14159 * deltaf + 7.95
        beta = 0 if A <= 21 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21) if A <= 50 else 0.1102 * (A - 8.7))
        f = f * kaiser_window(beta, length, device=device)
    f = f.unsqueeze(0)
    f = (f * f.T)[None, None]
    return f / f.sum()


def bilinear_filter(factor: int = 2, device="cpu") -> torch.Tensor:
    """(2 factor)^2 tent filter (functional/blur.py:339-366)"""
    if isinstance(factor, torch.Tensor):
        factor = factor.cpu().item()
    x = torch.arange(start=-factor + 0.5, end=factor, step=1, device=device) / factor
    w = 1 - x.abs()
    w = torch.outer(w, w)
    return (w / w.sum())[None, None]


def bicubic_filter(factor: int = 2, device="cpu") -> torch.Tensor:
    """(4 factor)^2 Keys cubic, a = -0.5 (functional/blur.py:552-583)"""
    if isinstance(factor, torch.Tensor):
        factor = factor.cpu().item()
    x = (torch.arange(start=-2 * factor + 0.5, end=2 * factor, step=1, device=device) / factor).abs()
    a = -0.5
    w = ((a + 2) * x.pow(3) - (a + 3) * x.pow(2) + 1) * (x <= 1)
    w = w + (a * x.pow(3) - 5 * a * x.pow(2) + 8 * a * x - 4 * a) * (x > 1) * (x < 2)
    w = torch.outer(w, w)
    return (w / w.sum())[None, None]

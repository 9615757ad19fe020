"""Denoising and Inpainting (deepinv/physics/forward.py:1255-1362, deepinv/physics/inpainting.py:12-185): the two
`DecomposablePhysics` whose singular vectors are the identity — the operators the diffusion samplers (DDRM, DiffPIR) are
usually demonstrated on.  A / A^T / prox are the generic SVD bodies (one elementwise pass over the mask); the samplers' spectral
updates and denoiser passes run on the libdinvk kernels as for MRI / BlurFFT."""
from __future__ import annotations

import torch

from .forward import DecomposablePhysics
from .noise import GaussianNoise


class Denoising(DecomposablePhysics):
    r"""y = x + noise (forward.py:1255-1362); Gaussian noise of standard deviation 0.1 unless a noise model is given"""

    def __init__(self, noise_model=None, device="cpu", **kwargs):
        if noise_model is None:
            noise_model = GaussianNoise(sigma=0.1)
        super().__init__(noise_model=noise_model, device=device, **kwargs)


class Inpainting(DecomposablePhysics):
    r"""y = m ⊙ x with a binary mask m broadcastable to the image (inpainting.py:12-185).  `mask`: a tensor, or a float p — then a
    Bernoulli(p) keep-mask is drawn once on `device` (per pixel, shared by the channels, when `pixelwise`; per entry otherwise)."""

    def __init__(self, img_size, mask=None, pixelwise: bool = True, device="cpu", rng: torch.Generator | None = None, **kwargs):
        super().__init__(device=device, **kwargs)
        if isinstance(mask, torch.Tensor):
            mask = mask.to(device)
        elif isinstance(mask, float):
            shape = (1, *img_size[-2:]) if (pixelwise and len(img_size) == 3) else tuple(img_size)
            keep = torch.rand(shape, device=device, generator=rng) < mask
            mask = keep.to(torch.float32).expand(*img_size).contiguous()
        elif mask is not None:
            raise ValueError("mask should either be torch.nn.Parameter, torch.Tensor, float or None.")
        if mask is not None and mask.dim() == len(img_size):
            mask = mask.unsqueeze(0)
        self.img_size = img_size
        self.register_buffer("mask", mask)
        self.to(device)

    def noise(self, x, **kwargs):
        return self.noise_model(x, **kwargs) * self.mask

    def __mul__(self, other):
        """masks of two inpainting (or an inpainting and an MRI) operators multiply (inpainting.py:150-178)"""
        from .mri import MRI

        if isinstance(other, Inpainting):
            return Inpainting(img_size=self.img_size, mask=self.mask * other.mask, noise_model=self.noise_model, device=self.mask.device)
        if isinstance(other, MRI):
            return other.__class__(mask=self.mask * other.mask, noise_model=self.noise_model, img_size=other.img_size,
                                   device=self.mask.device)
        return super().__mul__(other)

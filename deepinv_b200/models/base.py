"""Denoiser base class (deepinv/models/base.py:11-116): `forward(x, sigma)`."""
from __future__ import annotations

import torch
import torch.nn as nn


class Denoiser(nn.Module):
    def __init__(self, device=None):
        super().__init__()
        if device is not None:
            self.to(device)

    def forward(self, x: torch.Tensor, sigma, **kwargs) -> torch.Tensor:
        raise NotImplementedError

    @staticmethod
    def _handle_sigma(sigma, batch_size=None, ndim=None, device=None, dtype=None):
        """sigma as float / 0-d / (B,) -> (B,1,..,1) tensor (base.py:33-116)"""
        if not isinstance(sigma, torch.Tensor):
            sigma = torch.tensor(float(sigma), device=device, dtype=dtype)
        sigma = sigma.to(device=device, dtype=dtype)
        if sigma.dim() == 0:
            sigma = sigma.reshape(1)
        if batch_size is not None and sigma.numel() == 1:
            sigma = sigma.expand(batch_size)
        if ndim is not None:
            sigma = sigma.reshape((-1,) + (1,) * (ndim - 1))
        return sigma


def _no_grad_guard(name: str, *tensors) -> None:
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"{name}: the tensor-core (bf16) denoiser kernels are inference-only; train with precision='fp32' "
            "(differentiable: ops._ConvF32Fn) or run the bf16 denoiser under torch.no_grad()."
        )

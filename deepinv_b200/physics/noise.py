"""Noise models used to synthesise measurements (reference: deepinv/physics/noise.py:11-330).

Out of the hot path: plain torch RNG calls on the tensor's device, same call
semantics as the reference (`noise(x, sigma=None, seed=None)`, `update_parameters(sigma=...)`).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class NoiseModel(nn.Module):
    def __init__(self, noise_model=None, rng: torch.Generator | None = None):
        super().__init__()
        self.noise_model = (lambda x: x) if noise_model is None else noise_model
        self.rng = rng

    def forward(self, x, seed: int | None = None, **kwargs):
        self.rng_manual_seed(seed)
        return self.noise_model(x)

    def rng_manual_seed(self, seed: int | None = None):
        if seed is not None and self.rng is not None:
            self.rng = self.rng.manual_seed(seed)

    def randn_like(self, x, seed: int | None = None):
        self.rng_manual_seed(seed)
        return torch.empty_like(x).normal_(generator=self.rng)

    def update_parameters(self, **kwargs):
        for key, value in kwargs.items():
            if value is not None and hasattr(self, key) and isinstance(value, (torch.Tensor, float, int)):
                if not isinstance(value, torch.Tensor):
                    value = torch.tensor(float(value), dtype=torch.float32)
                self.register_buffer(key, value)


class ZeroNoise(NoiseModel):
    def forward(self, x, *args, **kwargs):
        return x


class GaussianNoise(NoiseModel):
    """y = x + sigma * n, n ~ N(0, I)  (noise.py:197-347)"""

    def __init__(self, sigma: float | torch.Tensor = 0.1, rng: torch.Generator | None = None):
        super().__init__(rng=rng)
        if not isinstance(sigma, torch.Tensor):
            sigma = torch.tensor(float(sigma), dtype=torch.float32)
        self.register_buffer("sigma", sigma)

    def forward(self, x, sigma=None, seed=None, **kwargs):
        self.update_parameters(sigma=sigma)
        sig = self.sigma.to(x.device)
        if sig.dim() > 0:
            sig = sig.reshape((sig.shape[0],) + (1,) * (x.dim() - 1))
        return x + self.randn_like(x, seed=seed) * sig

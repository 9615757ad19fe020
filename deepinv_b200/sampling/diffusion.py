"""DDRM sampler (drop-in for deepinv/sampling/diffusion.py:83-224).

Per step: x_bar = V^T x (one fused transform launch), the three-case spectral update + noise injection as ONE
elementwise kernel (`dinvk_ddrm_update`; the reference does ~15 boolean-indexed tensor ops, :196-217), then
x = denoiser(V x_bar, sigma_t).  Requires a DecomposablePhysics with a batch-1 mask, exactly like the
reference (:173).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..physics.forward import DecomposablePhysics


class DDRM(nn.Module):
    def __init__(self, denoiser, sigmas=None, eta: float = 0.85, etab: float = 1.0, verbose: bool = False, eps: float = 1e-6):
        super().__init__()
        if sigmas is None:
            sigmas = np.linspace(1, 0, 100)
        self.denoiser = denoiser
        self.sigmas = [float(s) for s in sigmas]
        self.max_iter = len(self.sigmas)
        self.eta, self.etab, self.verbose, self.eps = eta, etab, verbose, eps

    def forward(self, y, physics: DecomposablePhysics, seed=None, noises=None):
        """`noises` (optional list of tensors) replaces the successive torch.randn_like draws — used by the parity
        tests to feed the reference's exact noise sequence"""
        if not isinstance(physics, DecomposablePhysics):
            raise AttributeError(f"{type(physics).__name__} has no singular value decomposition (U_adjoint/V/V_adjoint); "
                                 "DDRM needs a DecomposablePhysics")
        with torch.no_grad():
            if seed:
                np.random.seed(seed)
                torch.manual_seed(seed)
            sigma_noise = float(physics.noise_model.sigma) if hasattr(physics.noise_model, "sigma") else 0.01
            mask = physics.mask.abs().to(torch.float32).contiguous()
            if mask.shape[0] != 1:
                raise IndexError("DDRM requires a batch-1 mask (deepinv/sampling/diffusion.py:173)")
            c = math.sqrt(1 - self.eta ** 2)
            y_bar = physics.U_adjoint(y).to(torch.float32).contiguous().clone()
            draw = (lambda t: next(it)) if noises is not None else (lambda t: torch.randn_like(y_bar))
            it = iter(noises) if noises is not None else None
            s = self.sigmas
            x_bar = ops.ddrm_update(None, None, y_bar, mask, draw(0).to(y_bar).contiguous(), s[0], 1.0, sigma_noise,
                                    self.eta, self.etab, 0.0, self.eps, init=True)
            x_bar_prev = x_bar
            x = self.denoiser(physics.V(x_bar), s[0])
            for t in range(1, self.max_iter):
                x_bar = physics.V_adjoint(x).contiguous()
                x_bar = ops.ddrm_update(x_bar, x_bar_prev, y_bar, mask, draw(t).to(y_bar).contiguous(), s[t], s[t - 1],
                                        sigma_noise, self.eta, self.etab, c * s[t], self.eps, init=False)
                x_bar_prev = x_bar
                x = self.denoiser(physics.V(x_bar), s[t])
        return x

"""L2 data fidelity (deepinv/optim/data_fidelity.py:237-338, optim/distance.py:82-95)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..physics.forward import LinearPhysics


class DataFidelity(nn.Module):
    def fn(self, x, y, physics, *args, **kwargs):
        raise NotImplementedError

    def forward(self, x, y, physics, *args, **kwargs):
        return self.fn(x, y, physics, *args, **kwargs)

    def prox_conjugate(self, x, y, physics, *args, gamma=1.0, lamb=1.0, **kwargs):
        """prox of (lamb f)^* by Moreau's identity (potential.py:120-133)"""
        return x - gamma * self.prox(x / gamma, y, physics, *args, gamma=lamb / gamma, **kwargs)


class ZeroFidelity(DataFidelity):
    def fn(self, x, y, physics, *args, **kwargs):
        return torch.zeros(x.shape[0], device=x.device)

    def grad(self, x, y, physics, *args, **kwargs):
        return torch.zeros_like(x)

    def prox(self, x, y, physics, *args, gamma=1.0, **kwargs):
        return x


class L2(DataFidelity):
    r"""f(x) = 1/(2 sigma^2) ||A x - y||^2"""

    def __init__(self, sigma: float = 1.0):
        super().__init__()
        self.sigma = sigma
        self.norm = 1 / (sigma ** 2)

    def d(self, u, y):
        diff = (u - y).reshape(u.shape[0], -1)
        return 0.5 * self.norm * (diff * diff).sum(-1)

    def fn(self, x, y, physics, *args, **kwargs):
        return self.d(physics.A(x), y)

    def grad(self, x, y, physics, *args, **kwargs):
        if isinstance(physics, LinearPhysics):
            return self.norm * (physics.A_adjoint_A(x) - physics.A_adjoint(y))
        return physics.A_vjp(x, self.norm * (physics.A(x) - y))

    def grad_d(self, u, y, *args, **kwargs):
        return self.norm * (u - y)

    def prox(self, x, y, physics, *args, gamma=1.0, **kwargs):
        return physics.prox_l2(x, y, self.norm * gamma)

    def prox_d(self, u, y, *args, gamma=1.0, **kwargs):
        g = self.norm * gamma
        return (u + g * y) / (1 + g)

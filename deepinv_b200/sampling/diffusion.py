"""DDRM and DiffPIR samplers (drop-ins for deepinv/sampling/diffusion.py:83-224 and :227-513).

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
            mask = physics.mask.abs().to(torch.float32)
            if mask.dim() == 0:  # Denoising: all singular values equal (diffusion.py:168-171 uses ones_like(y))
                mask = mask * torch.ones((1, *y.shape[1:]), dtype=torch.float32, device=y.device)
            mask = mask.contiguous()
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


class DiffPIR(nn.Module):
    r"""Diffusion plug-and-play restoration (diffusion.py:227-513, Zhu et al. 2023): per step a denoiser pass on the kernels, the
    data step `data_fidelity.prox` (closed-form spectral kernel for MRI / BlurFFT, CG on the operator kernels otherwise) and a
    DDIM-type resampling.

    The reference looks every per-step coefficient up in device-resident tables with device-side indices (`find_nearest`,
    `.argmin()`: several host synchronisations per step).  Here the whole schedule — noise level, the matching training
    timestep, alpha-bar, rho, the resampling weights — is a HOST plan of Python floats built once per noise level, so the loop
    issues kernels only."""

    def __init__(self, model, data_fidelity, sigma: float = 0.05, max_iter: int = 100, zeta: float = 0.1, lambda_: float = 7.0,
                 verbose: bool = False, device="cpu"):
        super().__init__()
        self.model, self.data_fidelity = model, data_fidelity
        self.sigma, self.max_iter, self.zeta, self.lambda_, self.verbose, self.device = sigma, max_iter, zeta, lambda_, verbose, device
        self.beta_start, self.beta_end, self.num_train_timesteps = 0.1 / 1000, 20 / 1000, 1000
        self._plan_sigma = None
        self._plan = None

    # ---- schedule (fp32 on the host, the reference's formulas :323-381 evaluated once) ------------------------------
    def _tables(self):
        betas = torch.linspace(self.beta_start, self.beta_end, self.num_train_timesteps, dtype=torch.float32)
        ac = torch.cumprod(1.0 - betas, dim=0)
        sqrt_ac, sqrt_1m = torch.sqrt(ac), torch.sqrt(1.0 - ac)
        return sqrt_ac, sqrt_1m, sqrt_1m / sqrt_ac, torch.sqrt(1.0 / ac)

    def make_plan(self, sigma: float) -> list[dict]:
        sqrt_ac, sqrt_1m, reduced, sqrt_recip = self._tables()
        T = self.num_train_timesteps
        sigmas = torch.flip(reduced, dims=(0,))                    # sigmas[i] = reduced[T-1-i]
        rhos = self.lambda_ * (sigma ** 2) / (sqrt_1m / sqrt_ac) ** 2
        seq = torch.sqrt(torch.linspace(0.0, T ** 2, self.max_iter)).to(torch.int32)
        seq[-1] = seq[-1] - 1
        nearest = lambda v: int(torch.abs(reduced - v).argmin())
        plan = []
        for i in range(len(seq)):
            cur = sigmas[seq[i]]
            t_i = nearest(cur)
            step = {"sigma": float(cur), "at": float(1 / sqrt_recip[t_i] ** 2), "last": bool(seq[i] == seq[-1]),
                    "sac_t": float(sqrt_ac[t_i]), "s1m_t": float(sqrt_1m[t_i]), "gamma": float(1.0 / (2 * rhos[t_i]))}
            if not step["last"]:
                t_n = nearest(sigmas[seq[i + 1]])
                step.update(sac_n=float(sqrt_ac[t_n]), s1m_n=float(sqrt_1m[t_n]))
            plan.append(step)
        # the INITIAL noise level uses the constructor's sigma, like the reference (diffusion.py:469); only rho / the noise
        # schedule follow the physics' noise model
        plan[0]["init_std"] = float((sigmas[seq[0]] ** 2 - 4.0 * float(self.sigma) ** 2).sqrt())
        plan[0]["init_div"] = float(sqrt_recip[-1])
        return plan

    def forward(self, y, physics, seed=None, x_init=None, noises=None):
        """`noises` (optional list of tensors) replaces the successive torch.randn_like draws (parity tests)"""
        if seed:
            torch.manual_seed(seed)
        sigma = float(physics.noise_model.sigma) if hasattr(physics.noise_model, "sigma") else float(self.sigma)
        if self._plan is None or self._plan_sigma != sigma:
            self._plan, self._plan_sigma = self.make_plan(sigma), sigma
        it = iter(noises) if noises is not None else None
        draw = (lambda t: next(it).to(t)) if noises is not None else torch.randn_like
        with torch.no_grad():
            x = 2 * (physics.A_adjoint(y) if x_init is None else x_init) - 1
            for i, st in enumerate(self._plan):
                if i == 0:
                    x = (x + st["init_std"] * draw(x)) / st["init_div"]
                if st["last"]:
                    continue  # the reference denoises once more here but never uses the result (:244-258, :267): skipped
                x_aux = x / (2 * st["at"] ** 0.5) + 0.5                       # to [0, 1] for the denoiser
                x0 = (2 * self.model(x_aux, st["sigma"] / 2) - 1).clamp(-1, 1)
                x0 = 2 * self.data_fidelity.prox(x0 / 2 + 0.5, y, physics, gamma=st["gamma"]) - 1
                eps = (x - st["sac_t"] * x0) / st["s1m_t"]                    # effective noise
                x = (st["sac_n"] * x0 + st["s1m_n"] * (1 - self.zeta) ** 0.5 * eps
                     + st["s1m_n"] * self.zeta ** 0.5 * draw(x))
        return x / 2 + 0.5

"""Least-squares solvers for non-decomposable operators (Tomography, Blur, MultiCoilMRI).

Mirror of deepinv/optim/linear/least_squares.py:15-197 (dispatcher) and
deepinv/optim/linear/conjugate_gradient.py:7-77 (CG).  The CG recurrences are the reference's
(same eps, same alpha/beta formulas, same "all samples below tol" stopping rule), but:
  * the three batched dot products, the scalar updates and the x/r/p updates are libdinvk kernels
    (`dinvk_batched_dot`, `dinvk_cg_scalars`, `dinvk_batched_axpy`);
  * the stopping test lives on the device (an int flag AND-ed over the batch); the host reads it
    every `check_every` iterations instead of synchronising on `torch.all(...)` every iteration
    (conjugate_gradient.py:61).  Iterations issued after convergence are frozen by the flag, so the
    result is the one the reference returns at its break.
"""
from __future__ import annotations

from typing import Callable

import torch

from .. import ops


def _dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return ops.batched_dot(a, b)


def conjugate_gradient(A: Callable, b: torch.Tensor, max_iter: int = 100, tol: float = 1e-5, eps: float = 1e-8,
                       init: torch.Tensor | None = None, verbose: bool = False, check_every: int = 1) -> torch.Tensor:
    """Solve A x = b for a symmetric positive operator, batch dimension 0 in parallel."""
    x = torch.zeros_like(b) if init is None else init
    r = ops.axpbypcz(b, 1.0, A(x), -1.0)
    p = r
    res_old = _dot(r, r)
    b_norm_sq = _dot(b, b)
    tol2 = float(tol) ** 2
    flag = torch.zeros(1, dtype=torch.int32, device=b.device)
    for i in range(int(max_iter)):
        Ap = A(p)
        alpha = ops.cg_scalars(0, res_old, _dot(p, Ap), eps, done_flag=flag)
        x = ops.batched_axpy(x, p, alpha, 1.0)
        r = ops.batched_axpy(r, Ap, alpha, -1.0)
        res_new = _dot(r, r)
        beta = ops.cg_scalars(1, res_new, res_old, eps, bnorm2=b_norm_sq, tol2=tol2, done_flag=flag)
        if (i % check_every == check_every - 1 or i == int(max_iter) - 1) and bool(flag.item()):
            if verbose:
                print("CG Converged at iteration", i + 1)
            break
        p = ops.batched_axpy(r, p, beta, 1.0)
        res_old = res_new
        if i > 0 and i % 100 == 0:
            r = ops.axpbypcz(b, 1.0, A(x), -1.0)
            res_old = _dot(r, r)
    else:
        if verbose:
            print("CG did not converge")
    return x


def bicgstab(A: Callable, b: torch.Tensor, init: torch.Tensor | None = None, max_iter: int = 100, tol: float = 1e-5,
             verbose: bool = False) -> torch.Tensor:
    """Stabilised bi-conjugate gradients for a square (not necessarily symmetric) operator, batch dimension in parallel
    (deepinv/optim/linear/bicgstab.py:8-107, van der Vorst 1992; same breakdown safeguards and stopping rule).  The seven
    inner products per iteration are `dinvk_batched_dot` launches, the vector updates `dinvk_batched_axpy` with the (B,) scalars
    staying on the device; only the convergence test reads one boolean back per iteration."""
    x = torch.zeros_like(b) if init is None else init
    r = ops.axpbypcz(b, 1.0, A(x), -1.0)
    r_hat = r.clone()
    rho = _dot(r, r_hat)
    p = r
    tol2 = _dot(b, b) * (float(tol) ** 2)
    eps = torch.finfo(torch.float32).eps
    safe = lambda num, den: torch.where(den.abs() > eps, num / den, torch.zeros_like(num))
    for i in range(int(max_iter)):
        v = A(p)
        alpha = safe(rho, _dot(r_hat, v))
        h = ops.batched_axpy(x, p, alpha, 1.0)
        s_ = ops.batched_axpy(r, v, alpha, -1.0)
        t = A(s_)
        omega = safe(_dot(t, s_), _dot(t, t))
        x = ops.batched_axpy(h, s_, omega, 1.0)
        r = ops.batched_axpy(s_, t, omega, -1.0)
        if bool(torch.all(_dot(r, r) < tol2)):
            if verbose:
                print("BiCGSTAB Converged at iteration", i)
            break
        rho_new = _dot(r, r_hat)
        ok = (rho.abs() > eps) & (omega.abs() > eps)
        beta = torch.where(ok, (rho_new / rho) * (alpha / omega), torch.zeros_like(rho_new))
        p = ops.batched_axpy(r, ops.batched_axpy(p, v, omega, -1.0), beta, 1.0)
        rho = rho_new
    else:
        if verbose:
            print("BiCGSTAB did not converge")
    return x


def minres(A: Callable, b: torch.Tensor, init: torch.Tensor | None = None, max_iter: int = 100, tol: float = 1e-5,
           eps: float = 1e-25, verbose: bool = False) -> torch.Tensor:
    """MINRES (Paige & Saunders 1975) for a symmetric, possibly indefinite operator, batch dimension in parallel
    (deepinv/optim/linear/minres.py:9-173): Lanczos three-term recurrence + one Givens rotation per step, on the right-hand side
    normalised per sample; stops when the update is below `tol` relative to the solution, like the reference.  Vector work is
    `dinvk_batched_dot` / `dinvk_batched_axpy`; the (B,) recurrence scalars stay on the device."""
    bnorm = _dot(b, b).sqrt()
    zero_b = bnorm < 1e-10
    bnorm = torch.where(zero_b, torch.ones_like(bnorm), bnorm)
    inv = 1.0 / bnorm
    scale = lambda t, s_: ops.batched_axpy(t, t, s_ - 1.0, 1.0)
    bn = scale(b, inv)
    x = torch.zeros_like(b) if init is None else scale(init, inv)
    z_prev2 = torch.zeros_like(b)
    z_prev1 = ops.axpbypcz(bn, 1.0, A(x), -1.0)
    beta_prev = _dot(z_prev1, z_prev1).sqrt().clamp_min(eps)
    z_prev1 = scale(z_prev1, 1.0 / beta_prev)
    q = z_prev1
    one, zero = torch.ones_like(beta_prev), torch.zeros_like(beta_prev)
    cos2, sin2, cos1, sin1 = one, zero, one, zero
    s_prev2, s_prev1 = torch.zeros_like(b), torch.zeros_like(b)
    scale_prev = beta_prev
    for i in range(int(max_iter)):
        prod = A(q)
        alpha = _dot(prod, q)
        prod = ops.batched_axpy(ops.batched_axpy(prod, z_prev1, alpha, -1.0), z_prev2, beta_prev, -1.0)
        beta = _dot(prod, prod).sqrt().clamp_min(eps)
        prod = scale(prod, 1.0 / beta)
        # apply the two previous rotations to the new tridiagonal column, then the new one
        subsub = sin2 * beta_prev
        sub = cos2 * beta_prev
        diag = alpha * cos1 - sin1 * sub
        sub = sub * cos1 + sin1 * alpha
        radius = torch.sqrt(diag * diag + beta * beta)
        cos, sin = diag / radius, beta / radius
        diag = diag * cos + sin * beta
        scale_cur = -scale_prev * sin
        search = ops.batched_axpy(ops.batched_axpy(q, s_prev1, sub, -1.0), s_prev2, subsub, -1.0)
        search = scale(search, 1.0 / diag)
        step = scale_prev * cos
        x = ops.batched_axpy(x, search, step, 1.0)
        upd = _dot(search, search).sqrt() * step.abs()
        if float((upd / _dot(x, x).sqrt()).max()) < tol:
            if verbose:
                print("MINRES converged at iteration", i + 1)
            break
        z_prev2, z_prev1, q, beta_prev = z_prev1, prod, prod, beta
        cos2, cos1, sin2, sin1 = cos1, cos, sin1, sin
        s_prev2, s_prev1, scale_prev = s_prev1, search, scale_cur
    x = scale(x, torch.where(zero_b, torch.zeros_like(bnorm), bnorm))
    return x


def lsqr(A: Callable, AT: Callable, b: torch.Tensor, eta=0.0, x0: torch.Tensor | None = None, tol: float = 1e-6,
         max_iter: int = 100, verbose: bool = False) -> torch.Tensor:
    r"""LSQR (Paige & Saunders 1982) for min_x ||A x - b||^2 + eta ||x - x0||^2 on rectangular operators, batch dimension in
    parallel (deepinv/optim/linear/lsqr.py:5-230; that file adapts SciPy's lsqr).  Golub-Kahan bidiagonalisation with the damped
    Givens update; per iteration one `A`, one `A^T`, two norms (`dinvk_batched_dot`) and five `dinvk_batched_axpy`, the
    (B,)-sized rotation scalars stay on the device; one boolean is read back per iteration for the stopping rule
    ||r|| <= tol ||b||.  eta: float or (B,) tensor."""
    nrm = lambda t: _dot(t, t).sqrt()
    scale = lambda t, s_: ops.batched_axpy(t, t, s_ - 1.0, 1.0)          # t * s[b]
    B = b.shape[0]
    dev = b.device
    eta_t = (eta.reshape(-1).float() if isinstance(eta, torch.Tensor) else torch.full((B,), float(eta or 0.0), device=dev))
    if bool((eta_t < 0).any()):
        raise ValueError("Damping parameter eta must be non-negative. LSQR cannot be applied to problems with negative eta.")
    damp = eta_t.sqrt()
    bnorm = nrm(b)
    if x0 is None:
        x = torch.zeros_like(AT(b))
        u = b
    else:
        x = x0.clone()
        u = ops.axpbypcz(b, 1.0, A(x), -1.0)
    beta = nrm(u)
    safe_inv = lambda t: torch.where(t > 0, 1.0 / t, torch.zeros_like(t))
    u = scale(u, safe_inv(beta))
    v = AT(u)
    alpha = nrm(v)
    v = scale(v, safe_inv(alpha))
    w = v
    rhobar, phibar = alpha, beta
    if bool(((alpha * beta) == 0).all()):
        return x
    for itn in range(int(max_iter)):
        u = ops.batched_axpy(A(v), u, alpha, -1.0)
        beta = nrm(u)
        u = scale(u, safe_inv(beta))
        v = ops.batched_axpy(AT(u), v, beta, -1.0)
        alpha = nrm(v)
        v = scale(v, safe_inv(alpha))
        # every division is guarded per sample: a degenerate sample in the batch (b = 0 or A^T b = 0 with eta = 0, e.g. a
        # padded all-zero measurement) has rhobar1 = rho = 0 and must stay at its start value instead of turning into NaN
        # (the reference returns early for it, lsqr.py:150-160, and skips the damped rotation when eta == 0)
        rhobar1 = torch.sqrt(rhobar ** 2 + eta_t)
        ok1 = rhobar1 > 0
        cs1 = torch.where(ok1, rhobar / torch.where(ok1, rhobar1, torch.ones_like(rhobar1)), torch.ones_like(rhobar1))
        sn1 = torch.where(ok1, damp / torch.where(ok1, rhobar1, torch.ones_like(rhobar1)), torch.zeros_like(rhobar1))
        psi, phibar = sn1 * phibar, cs1 * phibar
        rho = torch.hypot(rhobar1, beta)
        ok = rho > 0
        rho_s = torch.where(ok, rho, torch.ones_like(rho))
        cs = torch.where(ok, rhobar1 / rho_s, torch.ones_like(rho))
        sn = torch.where(ok, beta / rho_s, torch.zeros_like(rho))
        theta, rhobar = sn * alpha, -cs * alpha
        phi, phibar = cs * phibar, sn * phibar
        x = ops.batched_axpy(x, w, torch.where(ok, phi / rho_s, torch.zeros_like(rho)), 1.0)
        w = ops.batched_axpy(v, w, torch.where(ok, theta / rho_s, torch.zeros_like(rho)), -1.0)
        if bool((torch.sqrt(phibar ** 2 + psi ** 2) <= tol * bnorm).all()):
            if verbose:
                print("LSQR converged at iteration", itn)
            break
    else:
        if verbose:
            print("LSQR did not converge")
    return x


def _solve_normal(physics, y, z, init, g, g_batch, max_iter, tol, verbose, kwargs, solver="CG"):
    """CG on (A^T A + I/gamma) x = A^T y + z/gamma; gamma: None, a host float `g`, or a (B,) device tensor `g_batch`"""
    b = physics.A_adjoint(y, **kwargs)
    if g_batch is not None:
        inv = torch.reciprocal(g_batch)
        if z is not None:
            b = ops.batched_axpy(b, z, inv, 1.0)
        H = lambda v: ops.batched_axpy(physics.A_adjoint_A(v, **kwargs), v, inv, 1.0)
    elif g is not None:
        if z is not None:
            b = ops.axpbypcz(b, 1.0, z, 1.0 / g)
        H = lambda v: ops.axpbypcz(physics.A_adjoint_A(v, **kwargs), 1.0, v, 1.0 / g)
    else:
        H = lambda v: physics.A_adjoint_A(v, **kwargs)
    if solver == "BiCGStab":
        return bicgstab(H, b, init=init, max_iter=max_iter, tol=tol, verbose=verbose)
    if solver == "minres":
        return minres(H, b, init=init, max_iter=max_iter, tol=tol, verbose=verbose)
    return conjugate_gradient(H, b, max_iter=max_iter, tol=tol, init=init, verbose=verbose)


class _LeastSquaresFn(torch.autograd.Function):
    """h(z, y, gamma) = argmin_x gamma/2 ||Ax - y||^2 + 1/2 ||x - z||^2 with O(1)-memory backward by implicit
    differentiation (deepinv/optim/linear/least_squares.py:200-342): with M = (A^T A + I/gamma)^-1 and v the incoming
    gradient, ONE more CG solve gives Mv, and  dL/dy = A Mv,  dL/dz = Mv / gamma,  dL/dgamma = <Mv, h - z> / gamma^2.
    Both solves run on the operator kernels; nothing of the forward iteration is stored."""

    @staticmethod
    def forward(ctx, physics, y, z, init, gamma, opts):
        g, g_batch = _split_gamma(gamma, y.shape[0])
        with torch.no_grad():
            h = _solve_normal(physics, y, z, init, g, g_batch, opts["max_iter"], opts["tol"], opts["verbose"], opts["kwargs"],
                              opts["solver"])
        ctx.physics, ctx.opts, ctx.g, ctx.gamma_shape = physics, opts, g, (gamma.shape if isinstance(gamma, torch.Tensor) else None)
        ctx.save_for_backward(h, y, z, g_batch)
        return h

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        h, y, z, g_batch = ctx.saved_tensors
        physics, opts, g = ctx.physics, ctx.opts, ctx.g
        # (A^T A + I/gamma) mv = grad_output   <=>   the forward problem with y = 0, z = gamma * grad_output
        if g_batch is not None:
            zz = ops.batched_axpy(torch.zeros_like(grad_output), grad_output, g_batch, 1.0)
        else:
            zz = ops.axpbypcz(grad_output, g)
        mv = _solve_normal(physics, torch.zeros_like(y), zz, None, g, g_batch, opts["max_iter"], opts["tol"], False, opts["kwargs"],
                           opts["solver"])
        need = ctx.needs_input_grad
        gy = physics.A(mv, **opts["kwargs"]) if need[1] else None
        gz = None
        if need[2]:
            gz = ops.batched_axpy(torch.zeros_like(mv), mv, torch.reciprocal(g_batch), 1.0) if g_batch is not None \
                else ops.axpbypcz(mv, 1.0 / g)
        ggam = None
        if need[4]:
            per = ops.batched_dot(mv, ops.axpbypcz(h, 1.0, z, -1.0))
            if g_batch is not None:
                ggam = (per / g_batch ** 2).reshape(ctx.gamma_shape)
            else:
                ggam = (per.sum() / g ** 2).reshape(ctx.gamma_shape)
        return None, gy, gz, None, ggam, None


def _split_gamma(gamma, batch):
    """-> (host float | None, (B,) device tensor | None)"""
    if gamma is None:
        return None, None
    if isinstance(gamma, torch.Tensor) and gamma.numel() > 1:
        if gamma.shape[0] != batch or gamma.numel() != batch:
            raise ValueError("If gamma is batched, its batch size must match the one of y.")
        return None, gamma.detach().reshape(-1).float()
    return float(gamma), None


def least_squares(physics, y: torch.Tensor, z: torch.Tensor | None = None, init: torch.Tensor | None = None,
                  gamma=None, solver: str = "CG", max_iter: int = 100, tol: float = 1e-6, verbose: bool = False,
                  **kwargs) -> torch.Tensor:
    r"""argmin_x gamma/2 ||A x - y||^2 + 1/2 ||x - z||^2 via the normal equations
    (A^T A + I/gamma) x = A^T y + z/gamma  (least_squares.py:148-151); gamma may be a scalar or one value per sample.
    When a gradient w.r.t. y, z or gamma is being tracked the result carries the implicit-differentiation backward of the
    reference's `least_squares_implicit_backward` (least_squares.py:345-469)."""
    if solver not in ("CG", "cg", "BiCGStab", "lsqr", "minres", None):
        raise ValueError(f"Solver {solver} not recognized. Choose between 'CG', 'lsqr', 'BiCGStab' and 'minres'.")
    solver = solver if solver in ("BiCGStab", "lsqr", "minres") else "CG"
    kwargs.pop("parallel_dim", None)
    tracked = torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in (y, z, gamma))
    if tracked and solver in ("lsqr", "minres"):
        solver = "CG"  # the implicit-differentiation backward needs the normal-equation solve
    if tracked:
        if z is None:
            z = torch.zeros_like(physics.A_adjoint(y.detach(), **kwargs))
        if gamma is None:
            gamma = 1e8  # "no regularisation" of a tracked solve = the reference's pseudo-inverse branch (forward.py:850-862)
        gam = gamma if isinstance(gamma, torch.Tensor) else torch.tensor(float(gamma), device=y.device)
        opts = {"max_iter": max_iter, "tol": tol, "verbose": verbose, "kwargs": kwargs, "solver": solver}
        return _LeastSquaresFn.apply(physics, y, z, init, gam, opts)
    y0 = y if isinstance(y, torch.Tensor) else y[0]  # stacked operators measure TensorLists (physics/combine.py)
    if solver == "lsqr":  # rectangular solver on (A, A^T) directly: eta = 1/gamma, x0 = z (least_squares.py:118-130)
        if not isinstance(y, torch.Tensor):
            raise NotImplementedError("deepinv_b200: lsqr on stacked (TensorList) measurements is not supported; use CG")
        eta = 0.0 if gamma is None else (1.0 / gamma.reshape(-1).float() if isinstance(gamma, torch.Tensor) and gamma.numel() > 1
                                         else 1.0 / float(gamma))
        return lsqr(lambda v: physics.A(v, **kwargs), lambda v: physics.A_adjoint(v, **kwargs), y, eta=eta, x0=z, tol=tol,
                    max_iter=max_iter, verbose=verbose)
    if solver in ("BiCGStab", "minres") and isinstance(y, torch.Tensor):
        # the reference hands a "complete" system (A^T y has the shape of y) to BiCGStab / MINRES as A x = y itself — gamma and z
        # do not enter (least_squares.py:131-134); mirrored so that these solvers return what the reference returns
        probe = physics.A_adjoint(y, **kwargs)
        if probe.shape == y.shape:
            fn = bicgstab if solver == "BiCGStab" else minres
            return fn(lambda v: physics.A(v, **kwargs), y, init=init, max_iter=max_iter, tol=tol, verbose=verbose)
    g, g_batch = _split_gamma(gamma, y0.shape[0])
    return _solve_normal(physics, y, z, init, g, g_batch, max_iter, tol, verbose, kwargs, solver)

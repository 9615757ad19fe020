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


def least_squares(physics, y: torch.Tensor, z: torch.Tensor | None = None, init: torch.Tensor | None = None,
                  gamma=None, solver: str = "CG", max_iter: int = 100, tol: float = 1e-6, verbose: bool = False,
                  **kwargs) -> torch.Tensor:
    r"""argmin_x gamma/2 ||A x - y||^2 + 1/2 ||x - z||^2 via the normal equations
    (A^T A + I/gamma) x = A^T y + z/gamma  (least_squares.py:148-151)."""
    if solver not in ("CG", "cg", None):
        raise NotImplementedError(f"deepinv_b200: solver {solver!r} is outside the accelerated path (CG only, SURVEY §8 a12)")
    if isinstance(gamma, torch.Tensor) and gamma.numel() > 1:
        raise NotImplementedError("per-sample gamma is not supported by the CG path")
    g = None if gamma is None else float(gamma)
    b = physics.A_adjoint(y, **kwargs)
    if g is not None and z is not None:
        b = ops.axpbypcz(b, 1.0, z, 1.0 / g)
    if g is None:
        H = lambda v: physics.A_adjoint_A(v, **kwargs)
    else:
        H = lambda v: ops.axpbypcz(physics.A_adjoint_A(v, **kwargs), 1.0, v, 1.0 / g)
    x = conjugate_gradient(H, b, max_iter=max_iter, tol=tol, init=init, verbose=verbose)
    return x

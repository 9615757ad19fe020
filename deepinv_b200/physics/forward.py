"""Operator base classes: Physics -> LinearPhysics -> DecomposablePhysics.

Host-side mirror of deepinv/physics/forward.py:19-1252 (same method names, argument meaning and
error behaviour); the arithmetic of the concrete operators lives in libdinvk's kernels.  What is
kept verbatim from the reference contract:
  * `forward(x) = sensor(noise(A(x)))` (:107-120), `A_dagger` semantics, `update/update_parameters`
    storing tensors passed as kwargs into same-named buffers cast to the buffer's device/dtype
    (:249-276), `A_vjp = A_adjoint` for linear operators (:535-547), `A_adjoint_A`, `A_A_adjoint`
    (:549-571), `adjointness_test` (:696-723), `compute_sqnorm` by power iteration (:660-694,
    functional/matrix.py:5-44), CG-based `prox_l2`/`A_dagger` for non-decomposable operators
    (:751-862) and the closed forms for decomposable ones (:1212-1252).
"""
from __future__ import annotations

import copy
import warnings
from typing import Callable

import torch
import torch.nn as nn

from .noise import GaussianNoise, NoiseModel, ZeroNoise  # noqa: F401


class _LinearFn(torch.autograd.Function):
    """y = op(x) for a linear op with a known adjoint; backward applies the adjoint kernel, and the
    adjoint's backward applies the op again (adjoint-of-adjoint = A, forward.py:1338-1360)."""

    @staticmethod
    def forward(ctx, x, op, adj):
        ctx.op, ctx.adj = op, adj
        with torch.no_grad():
            return op(x)

    @staticmethod
    def backward(ctx, g):
        return _LinearFn.apply(g.contiguous(), ctx.adj, ctx.op), None, None


def linear_apply(x: torch.Tensor, op: Callable, adj: Callable) -> torch.Tensor:
    """apply `op` through autograd only when a gradient can flow"""
    if torch.is_grad_enabled() and x.requires_grad:
        return _LinearFn.apply(x, op, adj)
    return op(x)


class TensorKey:
    """Identity + version of the tensors a derived quantity (compressed mask, spectrum, trig table, memoised A^T y) was
    computed from.  It holds STRONG references to them: a key made of `(data_ptr(), _version)` alone collides when a tensor
    is freed and the caching allocator hands its address to the next tensor of the same shape (e.g. `algo(y.clone(), physics)`
    in a loop) — the stale entry would then be served for new data."""

    __slots__ = ("refs", "vers", "extra")

    def __init__(self, *tensors, extra=()):
        self.refs = tensors
        self.vers = tuple(t._version for t in tensors)
        self.extra = extra

    def matches(self, *tensors, extra=()) -> bool:
        return (len(tensors) == len(self.refs) and all(a is b for a, b in zip(tensors, self.refs))
                and self.vers == tuple(t._version for t in tensors) and self.extra == extra)


def cache_hit(key, *tensors, extra=()) -> bool:
    return key is not None and key.matches(*tensors, extra=extra)


class Physics(nn.Module):
    r"""y = N(A(x))  (forward.py:19-351)"""

    def __init__(self, A: Callable = lambda x, **kwargs: x, noise_model: NoiseModel | None = None,
                 sensor_model: Callable = lambda x: x, solver: str = "gradient_descent", max_iter: int = 50,
                 tol: float = 1e-4, **kwargs):
        super().__init__()
        self.noise_model = ZeroNoise() if noise_model is None else noise_model
        self.sensor_model = sensor_model
        self.forw = A
        self.SVD = False
        self.max_iter = max_iter
        self.tol = tol
        self.solver = solver
        if len(kwargs) > 0:
            warnings.warn(f"Arguments {kwargs} are passed to {self.__class__.__name__} but are ignored.")

    def forward(self, x, **kwargs):
        return self.sensor(self.noise(self.A(x, **kwargs), **kwargs))

    def A(self, x, **kwargs):
        return self.forw(x, **kwargs)

    def sensor(self, x):
        return self.sensor_model(x)

    def set_noise_model(self, noise_model, **kwargs):
        self.noise_model = noise_model

    def noise(self, x, **kwargs):
        return self.noise_model(x, **kwargs)

    def A_vjp(self, x, v):
        _, vjpfunc = torch.func.vjp(self.A, x)
        return vjpfunc(v)[0]

    def update(self, **kwargs):
        self.update_parameters(**kwargs)
        if hasattr(self.noise_model, "update_parameters"):
            self.noise_model.update_parameters(**kwargs)

    def update_parameters(self, **kwargs):
        for key, value in kwargs.items():
            if value is not None and hasattr(self, key) and isinstance(value, torch.Tensor):
                cur = getattr(self, key)
                if isinstance(cur, torch.Tensor):
                    if value.device.type != cur.device.type:
                        warnings.warn(
                            f"The provided tensor for parameter '{key}' is on a different device ({value.device}) "
                            f"than the current parameter device ({cur.device}). The current device will be used.",
                            stacklevel=2,
                        )
                    setattr(self, key, value.to(cur))
                else:
                    setattr(self, key, value)
                self._parameter_changed(key)

    def _parameter_changed(self, key: str) -> None:
        """hook: concrete operators drop derived device-side state (compressed masks, spectra) here"""

    def clone(self):
        return copy.deepcopy(self)

    def __mul__(self, other):
        """A = self o other  (forward.py:73-88, 573-583); keeps self's noise and sensor models"""
        from .combine import compose

        if not isinstance(self, LinearPhysics) or not isinstance(other, LinearPhysics):
            warnings.warn("You are composing two physics objects. The resulting physics will not retain the original attributes. "
                          "You may instead retrieve attributes of the original physics by indexing the resulting physics.")
        return compose(other, self, max_iter=self.max_iter, tol=self.tol)

    def stack(self, other):
        """[self; other] with TensorList measurements (forward.py:90-107, 585-601)"""
        from .combine import stack

        return stack(self, other)

    def set_ls_solver(self, solver, max_iter=None, tol=None):
        if max_iter is not None:
            self.max_iter = max_iter
        if tol is not None:
            self.tol = tol
        self.solver = solver

    def A_dagger(self, y, x_init=None):
        if self.solver != "gradient_descent":
            raise NotImplementedError(f"Solver {self.solver} not implemented for A_dagger")
        if x_init is None:
            if not hasattr(self, "A_adjoint"):
                raise ValueError("x_init must be provided for gradient descent solver if the physics does not have A_adjoint defined.")
            x_init = self.A_adjoint(y)
        x = x_init
        lr = 1e-1
        for _ in range(self.max_iter):
            x = x - lr * self.A_vjp(x, self.A(x) - y)
            if torch.nn.functional.mse_loss(self.A(x), y) < self.tol:
                break
        return x.clone()


class LinearPhysics(Physics):
    r"""Linear operator with adjoint (forward.py:354-862)."""

    def __init__(self, A=lambda x, **kwargs: x, A_adjoint=None, img_size=None, noise_model=None,
                 sensor_model=lambda x: x, max_iter=50, tol=1e-4, solver="CG", implicit_backward_solver: bool = True,
                 device="cpu", **kwargs):
        super().__init__(A=A, noise_model=noise_model, sensor_model=sensor_model, max_iter=max_iter, solver=solver,
                         tol=tol, **kwargs)
        self.A_adj = A_adjoint
        self.img_size = img_size
        self.implicit_backward_solver = implicit_backward_solver
        self.register_buffer("_device_holder", torch.tensor(0.0, device=device), persistent=False)
        self.to(device)

    def A_adjoint(self, y, **kwargs):
        if self.A_adj is None:
            raise ValueError("A_adjoint is not defined for this LinearPhysics (pass A_adjoint= to the constructor).")
        return self.A_adj(y, **kwargs)

    def A_vjp(self, x, v):
        return self.A_adjoint(v)

    def A_A_adjoint(self, y, **kwargs):
        return self.A(self.A_adjoint(y, **kwargs), **kwargs)

    def A_adjoint_A(self, x, **kwargs):
        return self.A_adjoint(self.A(x, **kwargs), **kwargs)

    def compute_sqnorm(self, x0, *, max_iter: int = 100, tol: float = 1e-3, verbose: bool = True, **kwargs):
        """squared spectral norm by power iteration (functional/matrix.py:5-44)"""
        x = torch.randn_like(x0)
        x = x / torch.linalg.vector_norm(x)
        zold = torch.zeros((), device=x0.device)
        z = zold
        for it in range(max_iter):
            y = self.A_adjoint_A(x, **kwargs)
            z = torch.sum(x.conj() * y).real / torch.linalg.vector_norm(x) ** 2
            rel_var = torch.abs(z - zold)
            if rel_var < tol:
                if verbose:
                    print(f"Power iteration converged at iteration {it}, ||A^T A||_2={z.item():.2f}")
                break
            zold = z
            x = y / torch.linalg.vector_norm(y)
        else:
            warnings.warn("Power iteration: convergence not reached")
        return z

    def compute_norm(self, x0, max_iter=100, tol=1e-3, verbose=True, squared=True, **kwargs):
        sq = self.compute_sqnorm(x0, max_iter=max_iter, tol=tol, verbose=verbose, **kwargs)
        return sq if squared else sq.sqrt()

    def adjointness_test(self, u, **kwargs):
        Au = self.A(u, **kwargs)
        v = torch.randn_like(Au)
        Atv = self.A_adjoint(v, **kwargs)
        s1 = (v.conj() * Au).flatten().sum()
        s2 = (Atv * u.conj()).flatten().sum()
        return s1.conj() - s2

    # --- least squares based prox / pseudo-inverse (forward.py:751-862) --------------------------
    def prox_l2(self, z, y, gamma, solver="CG", max_iter=None, tol=None, verbose=False, **kwargs):
        from ..optim.linear import least_squares

        if max_iter is not None:
            self.max_iter = max_iter
        if tol is not None:
            self.tol = tol
        if solver is not None:
            self.solver = solver
        if z is None or isinstance(z, (float, int)):
            z = torch.full_like(self.A_adjoint(y), fill_value=0.0 if z is None else float(z))
        return least_squares(self, y, z=z, init=z, gamma=gamma, solver=self.solver, max_iter=self.max_iter,
                             tol=self.tol, verbose=verbose, **kwargs)

    def A_dagger(self, y, solver="CG", max_iter=None, tol=None, verbose=False, **kwargs):
        from ..optim.linear import least_squares

        if max_iter is not None:
            self.max_iter = max_iter
        if tol is not None:
            self.tol = tol
        if solver is not None:
            self.solver = solver
        # gamma = 1e8 approximates the pseudo-inverse exactly as the reference's implicit-backward branch (:850-862)
        return least_squares(self, y, z=None, init=None, gamma=1e8, solver=self.solver, max_iter=self.max_iter,
                             tol=self.tol, verbose=verbose, **kwargs)


class DecomposablePhysics(LinearPhysics):
    r"""A = U diag(s) V^T with fast U, V (forward.py:990-1252)."""

    def __init__(self, U=None, V_adjoint=None, img_size=None, U_adjoint=None, V=None, mask=1.0, device="cpu", **kwargs):
        super().__init__(device=device, **kwargs)
        if U is None and U_adjoint is not None:
            raise ValueError("U must be provided if U_adjoint is provided.")
        if V_adjoint is None and V is not None:
            raise ValueError("V_adjoint must be provided if V is provided.")
        self._V_adjoint = (lambda x: x) if V_adjoint is None else V_adjoint
        self._U = (lambda x: x) if U is None else U
        self._U_adjoint = (lambda x: x) if U is None else U_adjoint
        self._V = (lambda x: x) if V_adjoint is None else V
        mask = torch.tensor(mask) if not isinstance(mask, torch.Tensor) else mask
        self.img_size = img_size
        self.register_buffer("mask", mask)
        self.to(device)

    def U(self, x):
        return self._U(x)

    def V(self, x, **kwargs):
        return self._V(x)

    def U_adjoint(self, x, **kwargs):
        return self._U_adjoint(x)

    def V_adjoint(self, x):
        return self._V_adjoint(x)

    def A(self, x, mask=None, **kwargs):
        self.update_parameters(mask=mask, **kwargs)
        return self.U(self.mask * self.V_adjoint(x))

    def A_adjoint(self, y, mask=None, **kwargs):
        self.update_parameters(mask=mask, **kwargs)
        return self.V(torch.conj(self.mask) * self.U_adjoint(y))

    def A_A_adjoint(self, y, mask=None, **kwargs):
        self.update_parameters(mask=mask, **kwargs)
        return self.U(self.mask.conj() * self.mask * self.U_adjoint(y))

    def A_adjoint_A(self, x, mask=None, **kwargs):
        self.update_parameters(mask=mask, **kwargs)
        return self.V(self.mask.conj() * self.mask * self.V_adjoint(x))

    def prox_l2(self, z, y, gamma, **kwargs):
        b = self.A_adjoint(y) + 1 / gamma * z
        if isinstance(gamma, torch.Tensor) and gamma.dim() < self.mask.dim():
            gamma = gamma[(...,) + (None,) * (self.mask.dim() - gamma.dim())]
            gamma = gamma.to(device=self.mask.device, dtype=self.mask.real.dtype)
        scaling = torch.conj(self.mask) * self.mask + 1 / gamma
        return self.V(self.V_adjoint(b) / scaling)

    def A_dagger(self, y, mask=None, **kwargs):
        self.update_parameters(mask=mask, **kwargs)
        m = torch.where(self.mask > 1e-5, self.mask.reciprocal(), 0.0)
        return self.V(self.U_adjoint(y) * m)

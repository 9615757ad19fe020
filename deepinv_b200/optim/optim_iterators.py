"""Per-algorithm iteration steps (deepinv/optim/optim_iterators/{optim_iterator,pgd,admm,hqs}.py).

Step algebra is the reference's.  When the data term is a plain `L2` on a
physics object that exposes a fused kernel (`normal_step`, closed-form `prox_l2`) and no gradient is
being recorded, the data step is ONE kernel launch that also folds the axpy algebra
(x - gamma*(A^T A x - A^T y)); otherwise the generic composition runs on the operator kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .data_fidelity import L2


def _axpby(a: float, x: torch.Tensor, b: float, y: torch.Tensor) -> torch.Tensor:
    """a*x + b*y in one kernel (falls back to torch arithmetic when a graph must be recorded)"""
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad or isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor)):
        return a * x + b * y
    if isinstance(a, torch.Tensor):
        a = float(a)
    if isinstance(b, torch.Tensor):
        b = float(b)
    return ops.axpbypcz(x, float(a), y, float(b))


def _scalar(v):
    return isinstance(v, (int, float)) or (isinstance(v, torch.Tensor) and v.numel() == 1 and not v.requires_grad)


class fStep(nn.Module):
    def __init__(self, g_first=False, **kwargs):
        super().__init__()
        self.g_first = g_first


class gStep(nn.Module):
    def __init__(self, g_first=False, **kwargs):
        super().__init__()
        self.g_first = g_first


class OptimIterator(nn.Module):
    """x_{k+1} = relax( g_step(f_step(x_k)) )  (optim_iterator.py:13-132)"""

    def __init__(self, g_first: bool = False, cost_fn=None, has_cost: bool = True, **kwargs):
        super().__init__()
        self.g_first = g_first
        self.has_cost = has_cost
        self.cost_fn = cost_fn
        self.f_step = fStep(g_first=g_first)
        self.g_step = gStep(g_first=g_first)

    def relaxation_step(self, u, v, beta):
        if _scalar(beta) and float(beta) == 1.0:
            return u
        return _axpby(beta, u, 1 - beta, v)

    def _cost(self, x, cur_data_fidelity, cur_prior, cur_params, y, physics):
        if self.cost_fn is not None and self.has_cost and cur_data_fidelity is not None and cur_prior is not None:
            return self.cost_fn(x, cur_data_fidelity, cur_prior, cur_params, y, physics)
        return None

    def forward(self, X, cur_data_fidelity, cur_prior, cur_params, y, physics, *args, **kwargs):
        x_prev = X["est"][0]
        aty = X.get("aty")
        if not self.g_first:
            z = self.f_step(x_prev, cur_data_fidelity, cur_params, y, physics, aty=aty)
            x = self.g_step(z, cur_prior, cur_params)
        else:
            z = self.g_step(x_prev, cur_prior, cur_params)
            x = self.f_step(z, cur_data_fidelity, cur_params, y, physics, aty=aty)
        x = self.relaxation_step(x, x_prev, cur_params["beta"])
        return {"est": (x, z), "cost": self._cost(x, cur_data_fidelity, cur_prior, cur_params, y, physics), "aty": aty}


# ---- PGD / FISTA -------------------------------------------------------------------------------
class fStepPGD(fStep):
    def forward(self, x, cur_data_fidelity, cur_params, y, physics, aty=None):
        if not self.g_first:
            step = cur_params["stepsize"]
            fused = (type(cur_data_fidelity) is L2 and hasattr(physics, "normal_step") and _scalar(step)
                     and not (torch.is_grad_enabled() and x.requires_grad))
            if fused:
                if aty is None:
                    aty = physics.A_adjoint(y)
                return physics.normal_step(x, aty, float(step) * cur_data_fidelity.norm)
            grad = step * cur_data_fidelity.grad(x, y, physics)
            return x - grad
        return cur_data_fidelity.prox(x, y, physics, gamma=cur_params["stepsize"])


class gStepPGD(gStep):
    def forward(self, x, cur_prior, cur_params):
        if not self.g_first:
            return cur_prior.prox(x, cur_params["g_param"], gamma=cur_params["lambda"] * cur_params["stepsize"])
        grad = cur_params["lambda"] * cur_params["stepsize"] * cur_prior.grad(x, cur_params["g_param"])
        return x - grad


class PGDIteration(OptimIterator):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.g_step = gStepPGD(**kwargs)
        self.f_step = fStepPGD(**kwargs)


class FISTAIteration(OptimIterator):
    """pgd.py:35-108 — momentum (k+a-1)/(k+a), a from cur_params (default 3)"""

    def __init__(self, a=3, **kwargs):
        super().__init__(**kwargs)
        self.g_step = gStepPGD(**kwargs)
        self.f_step = fStepPGD(**kwargs)
        self.a = a

    def forward(self, X, cur_data_fidelity, cur_prior, cur_params, y, physics, *args, **kwargs):
        x_prev, z_prev = X["est"][0], X["est"][1]
        aty = X.get("aty")
        k = 0 if "it" not in X else X["it"]
        a = cur_params.get("a", self.a)
        alpha = (k + a - 1) / (k + a)
        if not self.g_first:
            z = self.f_step(z_prev, cur_data_fidelity, cur_params, y, physics, aty=aty)
            x = self.g_step(z, cur_prior, cur_params)
        else:
            z = self.g_step(z_prev, cur_prior, cur_params)
            x = self.f_step(z, cur_data_fidelity, cur_params, y, physics, aty=aty)
        z = _axpby(1.0 + alpha, x, -alpha, x_prev)
        return {"est": (x, z), "cost": self._cost(x, cur_data_fidelity, cur_prior, cur_params, y, physics), "it": k + 1,
                "aty": aty}


# ---- ADMM ----------------------------------------------------------------------------------------
class fStepADMM(fStep):
    def forward(self, x, z, cur_data_fidelity, cur_params, y, physics):
        p = _axpby(1.0, x, 1.0 if self.g_first else -1.0, z)
        return cur_data_fidelity.prox(p, y, physics, gamma=cur_params["stepsize"])


class gStepADMM(gStep):
    def forward(self, x, z, cur_prior, cur_params):
        p = _axpby(1.0, x, -1.0 if self.g_first else 1.0, z)
        return cur_prior.prox(p, cur_params["g_param"], gamma=cur_params["lambda"] * cur_params["stepsize"])


class ADMMIteration(OptimIterator):
    """admm.py:38-77: u = prox_f(x - z); x = prox_g(u + z); z = z + beta (u - x)"""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.g_step = gStepADMM(**kwargs)
        self.f_step = fStepADMM(**kwargs)

    def forward(self, X, cur_data_fidelity, cur_prior, cur_params, y, physics, *args, **kwargs):
        x, z = X["est"]
        if z.shape != x.shape:
            z = torch.zeros_like(x)
        if self.g_first:
            u = self.g_step(x, z, cur_prior, cur_params)
            x = self.f_step(u, z, cur_data_fidelity, cur_params, y, physics)
        else:
            u = self.f_step(x, z, cur_data_fidelity, cur_params, y, physics)
            x = self.g_step(u, z, cur_prior, cur_params)
        beta = cur_params["beta"]
        if _scalar(beta) and not (torch.is_grad_enabled() and (u.requires_grad or x.requires_grad or z.requires_grad)):
            z = ops.axpbypcz(z, 1.0, u, float(beta), x, -float(beta))
        else:
            z = z + beta * (u - x)
        return {"est": (x, z), "cost": self._cost(x, cur_data_fidelity, cur_prior, cur_params, y, physics),
                "aty": X.get("aty")}


# ---- HQS -----------------------------------------------------------------------------------------
class fStepHQS(fStep):
    def forward(self, x, cur_data_fidelity, cur_params, y, physics, aty=None):
        return cur_data_fidelity.prox(x, y, physics, gamma=cur_params["stepsize"])


class gStepHQS(gStep):
    def forward(self, x, cur_prior, cur_params):
        return cur_prior.prox(x, cur_params["g_param"], gamma=cur_params["lambda"] * cur_params["stepsize"])


class HQSIteration(OptimIterator):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.g_step = gStepHQS(**kwargs)
        self.f_step = fStepHQS(**kwargs)


# ---- DRS -----------------------------------------------------------------------------------------
class fStepDRS(fStep):
    """prox_{gamma f} of z (f first) or of the reflection 2x - z (g first)  (drs.py:76-108)"""

    def forward(self, x, z, cur_data_fidelity, cur_params, y, physics):
        p = _axpby(2.0, x, -1.0, z) if self.g_first else z
        return cur_data_fidelity.prox(p, y, physics, gamma=cur_params["stepsize"])


class gStepDRS(gStep):
    """prox_{gamma lambda g} of the reflection 2x - z (f first) or of z (g first)  (drs.py:111-146)"""

    def forward(self, x, z, cur_prior, cur_params):
        p = z if self.g_first else _axpby(2.0, x, -1.0, z)
        return cur_prior.prox(p, cur_params["g_param"], gamma=cur_params["lambda"] * cur_params["stepsize"])


class DRSIteration(OptimIterator):
    """Douglas-Rachford splitting (drs.py:12-73): u = prox_f(z); x = prox_g(2u - z); z += beta (x - u)"""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.g_step = gStepDRS(**kwargs)
        self.f_step = fStepDRS(**kwargs)

    def forward(self, X, cur_data_fidelity, cur_prior, cur_params, y, physics, *args, **kwargs):
        x, z = X["est"] if len(X["est"]) == 2 else (None, X["est"])
        if x is not None and z.shape != x.shape:
            z = x  # the "dual" variable of DRS lives in the primal space
        if self.g_first:
            u = self.g_step(x, z, cur_prior, cur_params)
            x = self.f_step(u, z, cur_data_fidelity, cur_params, y, physics)
        else:
            u = self.f_step(x, z, cur_data_fidelity, cur_params, y, physics)
            x = self.g_step(u, z, cur_prior, cur_params)
        beta = cur_params["beta"]
        if _scalar(beta) and not (torch.is_grad_enabled() and (u.requires_grad or x.requires_grad or z.requires_grad)):
            z = ops.axpbypcz(z, 1.0, x, float(beta), u, -float(beta))
        else:
            z = z + beta * (x - u)
        return {"est": (x, z), "cost": self._cost(x, cur_data_fidelity, cur_prior, cur_params, y, physics),
                "aty": X.get("aty")}


# ---- GD ------------------------------------------------------------------------------------------
class fStepGD(fStep):
    def forward(self, x, cur_data_fidelity, cur_params, y, physics, aty=None):
        return cur_data_fidelity.grad(x, y, physics)


class gStepGD(gStep):
    def forward(self, x, cur_prior, cur_params):
        return cur_params["lambda"] * cur_prior.grad(x, cur_params["g_param"])


class GDIteration(OptimIterator):
    """x <- x - gamma (grad f(x) + lambda grad g(x))  (gradient_descent.py:12-66).

    With a plain L2 data term on a physics that has the fused normal-step kernel the data part
    x - gamma*norm*(A^T A x - A^T y) is one launch; the prior gradient is then subtracted by one axpby."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.g_step = gStepGD(**kwargs)
        self.f_step = fStepGD(**kwargs)

    def forward(self, X, cur_data_fidelity, cur_prior, cur_params, y, physics, *args, **kwargs):
        x_prev = X["est"][0]
        aty = X.get("aty")
        step = cur_params["stepsize"]
        gg = self.g_step(x_prev, cur_prior, cur_params)
        fused = (type(cur_data_fidelity) is L2 and hasattr(physics, "normal_step") and _scalar(step) and aty is not None
                 and not (torch.is_grad_enabled() and (x_prev.requires_grad or gg.requires_grad)))
        if fused:
            x = _axpby(1.0, physics.normal_step(x_prev, aty, float(step) * cur_data_fidelity.norm), -float(step), gg)
        else:
            x = x_prev - step * (gg + self.f_step(x_prev, cur_data_fidelity, cur_params, y, physics))
        return {"est": (x,), "cost": self._cost(x, cur_data_fidelity, cur_prior, cur_params, y, physics), "aty": aty}


# ---- Chambolle-Pock (primal-dual) ----------------------------------------------------------------------
class fStepCP(fStep):
    def forward(self, x, w, cur_data_fidelity, y, physics, cur_params):
        if self.g_first:
            return cur_data_fidelity.prox(x - cur_params["stepsize"] * w, y, physics, gamma=cur_params["stepsize"])
        return cur_data_fidelity.prox_conjugate(x + cur_params["stepsize_dual"] * w, y, physics, gamma=cur_params["stepsize_dual"])


class gStepCP(gStep):
    def forward(self, x, w, cur_prior, cur_params):
        if self.g_first:
            return cur_prior.prox_conjugate(x + cur_params["stepsize_dual"] * w, cur_params["g_param"],
                                            gamma=cur_params["lambda"] * cur_params["stepsize_dual"], lamb=cur_params["lambda"])
        return cur_prior.prox(x - cur_params["stepsize"] * w, cur_params["g_param"],
                              gamma=cur_params["stepsize"] * cur_params["lambda"])


class CPIteration(OptimIterator):
    """Chambolle-Pock primal-dual iteration on (x, z, u) = (primal, extrapolated primal, dual) with an optional linear map K
    (primal_dual_CP.py:12-173):  u <- prox_{sigma F*}(u + sigma K z);  x <- prox_{tau G}(x - tau K^T u);  z <- x + beta (x - x_prev),
    F / G being the data term and the prior in the order chosen by `g_first`."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.g_step = gStepCP(**kwargs)
        self.f_step = fStepCP(**kwargs)

    def forward(self, X, cur_data_fidelity, cur_prior, cur_params, y, physics, *args, **kwargs):
        x_prev, z_prev, u_prev = X["est"]
        K = cur_params.get("K") or (lambda v: v)
        Kt = cur_params.get("K_adjoint") or (lambda v: v)
        if self.g_first:
            u = self.g_step(u_prev, K(z_prev), cur_prior, cur_params)
            x = self.f_step(x_prev, Kt(u), cur_data_fidelity, y, physics, cur_params)
        else:
            u = self.f_step(u_prev, K(z_prev), cur_data_fidelity, y, physics, cur_params)
            x = self.g_step(x_prev, Kt(u), cur_prior, cur_params)
        z = x + cur_params["beta"] * (x - x_prev)
        return {"est": (x, z, u), "cost": self._cost(x, cur_data_fidelity, cur_prior, cur_params, y, physics), "aty": X.get("aty")}

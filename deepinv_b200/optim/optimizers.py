"""Loop driver and named algorithms (deepinv/optim/optimizers.py:94-881, 1065-1734; fixed_point.py:262-406).

Same constructor keywords, parameter handling (`sigma_denoiser`->`g_param`, `lambda_reg`->`lambda`,
scalars or per-iteration lists, `unfold=True` turning them into nn.Parameters), initialisation
x0 = z0 = A^T y (computed ONCE here and carried in the iterate as "aty"; the reference evaluates it
twice at init and again inside every L2.grad), convergence criterion (batch-mean relative residual,
optimizers.py:703-739) and metrics dictionary as the reference.
"""
from __future__ import annotations

import warnings
from collections.abc import Iterable
from contextlib import nullcontext

import torch
import torch.nn as nn

from .data_fidelity import ZeroFidelity
from dataclasses import dataclass

from .optim_iterators import (ADMMIteration, CPIteration, DRSIteration, FISTAIteration, GDIteration, HQSIteration,
                              OptimIterator, PGDIteration)
from .prior import ZeroPrior


@dataclass
class AndersonAccelerationConfig:
    """Anderson acceleration of the fixed-point loop (optimizers.py:64-78)"""
    history_size: int = 10
    beta: float = 0.9
    eps: float = 0.1
    full_backprop: bool = False


@dataclass
class BacktrackingConfig:
    """Armijo-type backtracking on the stepsize (optimizers.py:80-91)"""
    gamma: float = 0.1
    eta: float = 0.9
    max_iter: int = 20


class _AndersonState:
    """Type-II Anderson mixing of the last m iterates (fixed_point.py:117-260): with G = T - X the residuals of the stored
    pairs, solve the bordered system [[0, 1^T], [1, G G^T + eps I]] [nu; p] = [1; 0] per sample (weights p sum to 1) and take
    x = beta * p^T T + (1 - beta) * p^T X.  The iterates are whatever the kernels produced; the mixing itself is two small
    batched GEMMs (m <= history_size rows) and an (m+1) x (m+1) solve per sample — library calls, not a hot path."""

    def __init__(self, cfg: AndersonAccelerationConfig, x: torch.Tensor):
        B, d, m = x.shape[0], x[0].numel(), cfg.history_size
        self.cfg = cfg
        self.x_hist = torch.zeros(B, m, d, dtype=x.dtype, device=x.device)
        self.t_hist = torch.zeros(B, m, d, dtype=x.dtype, device=x.device)
        self.H = torch.zeros(B, m + 1, m + 1, dtype=x.dtype, device=x.device)
        self.H[:, 0, 1:] = 1.0
        self.H[:, 1:, 0] = 1.0
        self.q = torch.zeros(B, m + 1, 1, dtype=x.dtype, device=x.device)
        self.q[:, 0] = 1.0

    def step(self, it: int, x_prev: torch.Tensor, tx: torch.Tensor) -> torch.Tensor:
        cfg = self.cfg
        B = x_prev.shape[0]
        slot, m = it % cfg.history_size, min(it + 1, cfg.history_size)
        self.x_hist[:, slot] = x_prev.reshape(B, -1).detach()
        self.t_hist[:, slot] = tx.reshape(B, -1).detach()
        X, T = self.x_hist[:, :m], self.t_hist[:, :m]
        if torch.is_grad_enabled() and (x_prev.requires_grad or tx.requires_grad):
            # gradient through the current pair only (full_backprop=False semantics, fixed_point.py:218-235)
            sel = torch.zeros(1, m, 1, dtype=X.dtype, device=X.device)
            sel[:, slot] = 1.0
            X = X + sel * (x_prev.reshape(B, 1, -1) - X[:, slot: slot + 1])
            T = T + sel * (tx.reshape(B, 1, -1) - T[:, slot: slot + 1])
        G = T - X
        H = self.H.clone()
        H[:, 1: m + 1, 1: m + 1] = torch.bmm(G, G.transpose(1, 2)) + cfg.eps * torch.eye(m, dtype=X.dtype, device=X.device)[None]
        p = torch.linalg.solve(H[:, : m + 1, : m + 1], self.q[:, : m + 1])[:, 1: m + 1, 0]
        self.H = H.detach()
        x = cfg.beta * (p[:, None] @ T)[:, 0] + (1 - cfg.beta) * (p[:, None] @ X)[:, 0]
        return x.view_as(x_prev)


@dataclass
class DEQConfig:
    """backward-pass settings of a deep-equilibrium model (optimizers.py:44-62)"""
    max_iter_backward: int = 50
    anderson_acceleration_backward: bool = False
    history_size_backward: int = 5
    beta_backward: float = 1.0
    eps_backward: float = 1e-4
    jacobian_free: bool = False


def objective_function(x, data_fidelity, prior, cur_params, y, physics):
    return data_fidelity(x, y, physics) + cur_params["lambda"] * prior(x, cur_params["g_param"])


class BaseOptim(nn.Module):
    def __init__(self, iterator: OptimIterator, params_algo=None, data_fidelity=None, prior=None, max_iter: int = 100,
                 crit_conv: str = "residual", thres_conv: float = 1e-5, early_stop: bool = False, has_cost: bool = False,
                 custom_metrics=None, custom_init=None, get_output=lambda X: X["est"][0], unfold: bool = False,
                 trainable_params=None, verbose: bool = False, show_progress_bar: bool = False, DEQ=None,
                 anderson_acceleration=False, backtracking=None, **kwargs):
        super().__init__()
        if isinstance(backtracking, bool):
            self.backtracking, self.backtracking_config = backtracking, (BacktrackingConfig() if backtracking else None)
        else:
            self.backtracking, self.backtracking_config = backtracking is not None, (backtracking or BacktrackingConfig())
        if isinstance(anderson_acceleration, bool):
            self.anderson_acceleration_config = AndersonAccelerationConfig() if anderson_acceleration else None
        else:
            self.anderson_acceleration_config = anderson_acceleration
        if self.anderson_acceleration_config is not None and self.anderson_acceleration_config.full_backprop:
            raise NotImplementedError("deepinv_b200: Anderson acceleration with full_backprop=True is not supported")
        self._anderson = None
        if isinstance(DEQ, bool):
            self.DEQ, self.DEQ_config = DEQ, (DEQConfig() if DEQ else None)
        else:
            self.DEQ, self.DEQ_config = DEQ is not None, (DEQ or DEQConfig())
        if self.DEQ and self.DEQ_config.anderson_acceleration_backward:
            raise NotImplementedError("deepinv_b200: Anderson-accelerated DEQ backward is outside the accelerated path")
        self.early_stop, self.crit_conv, self.verbose = early_stop, crit_conv, verbose
        self.max_iter, self.thres_conv = max_iter, thres_conv
        self.custom_metrics, self.custom_init, self.get_output = custom_metrics, custom_init, get_output
        self.unfold = unfold
        self.has_converged = False
        self.iterator = iterator
        self.prior = [ZeroPrior()] if prior is None else (list(prior) if isinstance(prior, Iterable) else [prior])
        self.data_fidelity = [ZeroFidelity()] if data_fidelity is None else (
            list(data_fidelity) if isinstance(data_fidelity, Iterable) else [data_fidelity])
        self.has_cost = self.prior[0].explicit_prior
        iterator.has_cost = self.has_cost
        if self.has_cost and iterator.cost_fn is None:
            iterator.cost_fn = objective_function

        params_algo = dict({"lambda": 1.0, "stepsize": 1.0} if params_algo is None else params_algo)
        if "g_param" not in params_algo:
            params_algo["g_param"] = params_algo.pop("sigma_denoiser", None)
        if "lambda" not in params_algo:
            params_algo["lambda"] = params_algo.pop("lambda_reg", 1.0)
        params_algo.setdefault("beta", 1.0)
        for key, value in params_algo.items():
            if not isinstance(value, Iterable) or (isinstance(value, torch.Tensor) and value.dim() == 0):
                params_algo[key] = [value]
            elif 1 < len(value) < self.max_iter:
                raise ValueError(f"The number of elements in the parameter {key} is inferior to max_iter.")
        self.init_params_algo = params_algo
        self._host_schedules = {}
        if self.backtracking and len(params_algo["stepsize"]) > 1:
            warnings.warn("Backtracking impossible when stepsize is predefined as a list. Setting backtracking to False.")
            self.backtracking = False
        if self.backtracking and not self.has_cost:
            warnings.warn("Backtracking impossible when no cost function is given. Setting backtracking to False.")
            self.backtracking = False

        if self.unfold or self.DEQ:
            if trainable_params is not None:
                trainable_params = ["lambda" if p == "lambda_reg" else "g_param" if p == "sigma_denoiser" else p
                                    for p in trainable_params]
            else:
                trainable_params = list(params_algo.keys())
            for k in trainable_params:
                if k in self.init_params_algo and self.init_params_algo[k][0] is not None:
                    self.init_params_algo[k] = nn.ParameterList([
                        nn.Parameter(torch.tensor(el).float() if not isinstance(el, torch.Tensor) else el.float())
                        for el in self.init_params_algo[k]])
            self.params_algo = nn.ParameterDict(
                {k: v for k, v in self.init_params_algo.items() if isinstance(v, nn.ParameterList)})
            self.prior = nn.ModuleList(self.prior)
            self.data_fidelity = nn.ModuleList(self.data_fidelity)

    # ---- per-iteration lookups (optimizers.py:464-502) --------------------------------------------
    def update_params_fn(self, it: int) -> dict:
        """parameters of iteration `it` (optimizers.py:464-480).  Fixed (non-trainable) schedules given as tensors — e.g.
        DPIR's per-iteration sigma / stepsize on the device — are read back to host floats ONCE, so that the loop does not
        synchronise on a device scalar every iteration (the values are the same fp32 numbers)."""
        out = {}
        for k, v in self.init_params_algo.items():
            if isinstance(v, torch.Tensor) and not v.requires_grad and v.dim() == 1:
                host = self._host_schedules.get(k)
                if host is None or host[0] is not v:
                    host = (v, v.detach().cpu().tolist())
                    self._host_schedules[k] = host
                v = host[1]
            out[k] = v[it] if len(v) > 1 else v[0]
        return out

    def update_prior_fn(self, it: int):
        return self.prior[it] if len(self.prior) > 1 else self.prior[0]

    def update_data_fidelity_fn(self, it: int):
        return self.data_fidelity[it] if len(self.data_fidelity) > 1 else self.data_fidelity[0]

    def init_iterate_fn(self, y, physics, init=None):
        init = init if init is not None else self.custom_init
        if init is not None:
            if callable(init):
                init = init(y, physics)
            if isinstance(init, torch.Tensor):
                X = {"est": (init,)}
            elif isinstance(init, tuple):
                X = {"est": init}
            elif isinstance(init, dict):
                X = dict(init)
            else:
                raise ValueError(f"Custom initial iterate must be a torch.Tensor, a tuple, or a dict. Got {type(init)}.")
            if len(X["est"]) == 1:
                X["est"] = (X["est"][0], X["est"][0])
            X.setdefault("aty", None)
        else:
            aty = physics.A_adjoint(y)
            X = {"est": (aty, aty.clone()), "aty": aty}
        X["cost"] = None
        if self.has_cost and self.iterator.cost_fn is not None:  # F(x0): backtracking / cost-based stopping compare against it
            X["cost"] = self.iterator.cost_fn(X["est"][0], self.update_data_fidelity_fn(0), self.update_prior_fn(0),
                                              self.update_params_fn(0), y, physics)
        return X

    def backtracking_check_fn(self, X_prev, X) -> bool:
        """sufficient-decrease test F(x_prev) - F(x) >= gamma / stepsize * ||x - x_prev||^2 (batch means); on failure the
        stepsize is multiplied by eta and the iterate is rejected (optimizers.py:668-701)"""
        if not (self.backtracking and self.has_cost and X_prev is not None and X_prev.get("cost") is not None):
            return True
        x_prev = X_prev["est"][0].reshape(X_prev["est"][0].shape[0], -1)
        x = X["est"][0].reshape(x_prev.shape[0], -1)
        diff_F = (X_prev["cost"] - X["cost"]).mean()
        diff_x = torch.linalg.vector_norm(x - x_prev, dim=-1, ord=2).pow(2).mean()
        stepsize = self.init_params_algo["stepsize"][0]
        if diff_F < (self.backtracking_config.gamma / stepsize) * diff_x:
            self.init_params_algo["stepsize"] = [self.backtracking_config.eta * stepsize]
            if self.verbose:
                print(f"Backtracking : new stepsize = {float(self.init_params_algo['stepsize'][0]):.6f}")
            return False
        return True

    def check_conv_fn(self, it, X_prev, X) -> bool:
        if self.crit_conv == "residual":
            x_prev = self.get_output(X_prev).reshape(self.get_output(X_prev).shape[0], -1)
            x = self.get_output(X).reshape(x_prev.shape[0], -1)
            crit_cur = ((x_prev - x).norm(p=2, dim=-1) / (x.norm(p=2, dim=-1) + 1e-06)).mean()
        elif self.crit_conv == "cost":
            F_prev, F = X_prev["cost"], X["cost"]
            crit_cur = ((F_prev - F).norm(dim=-1) / (F.norm(dim=-1) + 1e-06)).mean()
        else:
            raise ValueError("convergence criteria not implemented")
        if crit_cur < self.thres_conv:
            self.has_converged = True
            if self.verbose:
                print(f"Iteration {it}, current converge crit. = {crit_cur:.2E}, objective = {self.thres_conv:.2E}")
            return True
        return False

    def _metrics_init(self, X, x_gt):
        x0 = self.get_output(X)
        self.batch_size = x0.shape[0]
        m = {"psnr": [[] for _ in range(self.batch_size)], "residual": [[] for _ in range(self.batch_size)]}
        if x_gt is not None:
            for i in range(self.batch_size):
                m["psnr"][i].append(_psnr(x0[i:i + 1], x_gt[i:i + 1]))
        if self.has_cost:
            m["cost"] = [[] for _ in range(self.batch_size)]
        if self.custom_metrics is not None:
            for name in self.custom_metrics:
                m[name] = [[] for _ in range(self.batch_size)]
        return m

    def _metrics_update(self, m, X_prev, X, x_gt):
        x_prev, x = self.get_output(X_prev), self.get_output(X)
        for i in range(self.batch_size):
            m["residual"][i].append(((x_prev[i] - x[i]).norm() / (x[i].norm() + 1e-06)).detach().cpu().item())
            if x_gt is not None:
                m["psnr"][i].append(_psnr(x[i:i + 1], x_gt[i:i + 1]))
            if self.has_cost:
                m["cost"][i].append(X["cost"][i].detach().cpu().item())
            if self.custom_metrics is not None:
                for name, fn in self.custom_metrics.items():
                    m[name][i].append(fn(m[name], x_prev[i], x[i]))
        return m

    def single_iteration(self, X, it, y, physics, **kwargs):
        fid, prior, params = self.update_data_fidelity_fn(it), self.update_prior_fn(it), self.update_params_fn(it)
        Xn = self.iterator(X, fid, prior, params, y, physics, **kwargs)
        if self.anderson_acceleration_config is not None:  # mix the new iterate with the history (fixed_point.py:391-400)
            if self._anderson is None or it == 0:
                self._anderson = _AndersonState(self.anderson_acceleration_config, X["est"][0])
            x = self._anderson.step(it, X["est"][0], Xn["est"][0])
            Xn = dict(Xn)
            Xn["est"] = (x, *Xn["est"][1:])
            if self.has_cost and self.iterator.cost_fn is not None:
                Xn["cost"] = self.iterator.cost_fn(x, fid, prior, params, y, physics)
        return Xn

    def DEQ_additional_step(self, X, y, physics, **kwargs):
        """One more iteration WITH gradient tracking at the equilibrium x*, plus a backward hook that replaces the
        incoming gradient u by the solution of v = J(x*)^T v + u, found by `max_iter_backward` fixed-point sweeps
        (optimizers.py:741-828; implicit function theorem).  Every J^T v is one autograd pass through the iteration,
        i.e. the adjoint operator kernels and the denoiser's backward kernels."""
        it = self.max_iter - 1
        fid, prior, params = self.update_data_fidelity_fn(it), self.update_prior_fn(it), self.update_params_fn(it)
        x = self.iterator(X, fid, prior, params, y, physics, **kwargs)["est"][0]
        if not self.DEQ_config.jacobian_free and x.requires_grad:
            x0 = x.detach().clone().requires_grad_()
            f0 = self.iterator({"est": (x0,), "aty": X.get("aty")}, fid, prior, params, y, physics, **kwargs)["est"][0]
            n_back = int(self.DEQ_config.max_iter_backward)

            def solve_adjoint_fixed_point(grad):
                v = grad
                for _ in range(n_back):
                    v = torch.autograd.grad(f0, x0, v, retain_graph=True)[0] + grad
                return v

            x.register_hook(solve_adjoint_fixed_point)
        return x

    def forward(self, y, physics, init=None, x_gt=None, compute_metrics: bool = False, **kwargs):
        ctx = torch.no_grad() if (not self.unfold or self.DEQ) else nullcontext()
        with ctx:
            X = self.init_iterate_fn(y, physics, init=init)
            metrics = self._metrics_init(X, x_gt) if compute_metrics else None
            self.has_converged = False
            failed = 0
            for it in range(self.max_iter):
                X_prev = X
                X = self.single_iteration(X, it, y, physics, **kwargs)
                if not self.backtracking_check_fn(X_prev, X):  # rejected step: keep the iterate, retry with the smaller stepsize
                    X = X_prev
                    failed += 1
                    if failed >= self.backtracking_config.max_iter:
                        break
                    continue
                failed = 0
                if compute_metrics:
                    metrics = self._metrics_update(metrics, X_prev, X, x_gt)
                if self.early_stop and it > 1 and self.check_conv_fn(it, X_prev, X):
                    break
        x = self.DEQ_additional_step(X, y, physics, **kwargs) if self.DEQ else self.get_output(X)
        return (x, metrics) if compute_metrics else x


def _psnr(x, y, max_pixel=1.0):
    mse = ((x - y) ** 2).mean()
    return float(10 * torch.log10(max_pixel ** 2 / mse))


def _named(iteration_cls):
    class _Algo(BaseOptim):
        def __init__(self, data_fidelity=None, prior=None, lambda_reg: float = 1.0, stepsize: float = 1.0,
                     g_param=None, sigma_denoiser=None, beta: float = 1.0, max_iter: int = 100,
                     crit_conv: str = "residual", thres_conv: float = 1e-5, early_stop: bool = False,
                     custom_metrics=None, custom_init=None, g_first: bool = False, unfold: bool = False,
                     trainable_params=None, cost_fn=None, params_algo=None, **kwargs):
            if g_param is None and sigma_denoiser is not None:
                g_param = sigma_denoiser
            if params_algo is None:
                params_algo = {"lambda": lambda_reg, "stepsize": stepsize, "g_param": g_param, "beta": beta}
            super().__init__(iteration_cls(g_first=g_first, cost_fn=cost_fn), data_fidelity=data_fidelity, prior=prior,
                             params_algo=params_algo, max_iter=max_iter, crit_conv=crit_conv, thres_conv=thres_conv,
                             early_stop=early_stop, custom_metrics=custom_metrics, custom_init=custom_init,
                             unfold=unfold, trainable_params=trainable_params, **kwargs)

    return _Algo


class PGD(_named(PGDIteration)):
    """Proximal gradient descent (optimizers.py:1596-1734)"""


class FISTA(_named(FISTAIteration)):
    """FISTA (optimizers.py:1737-1880)"""

    def __init__(self, *args, a: int = 3, **kwargs):
        super().__init__(*args, **kwargs)
        self.init_params_algo.setdefault("a", [a])


class ADMM(_named(ADMMIteration)):
    """ADMM (optimizers.py:1065-1191)"""


class HQS(_named(HQSIteration)):
    """Half-quadratic splitting (optimizers.py:1459-1593)"""


class DRS(_named(DRSIteration)):
    """Douglas-Rachford splitting (optimizers.py:1194-1317)"""


class GD(_named(GDIteration)):
    """Gradient descent on f + lambda g (optimizers.py:1320-1456)"""


class PDCP(BaseOptim):
    """Primal-dual Chambolle-Pock (optimizers.py:2088-2245): iterates (x, z, u) initialised as (A^T y, A^T y, y)"""

    def __init__(self, data_fidelity=None, prior=None, lambda_reg: float = 1.0, stepsize: float = 1.0, stepsize_dual: float = 1.0,
                 beta: float = 1.0, K=None, K_adjoint=None, g_param=None, sigma_denoiser=None, max_iter: int = 100,
                 crit_conv: str = "residual", thres_conv: float = 1e-5, early_stop: bool = False, custom_metrics=None,
                 custom_init=None, g_first: bool = False, unfold: bool = False, trainable_params=None, cost_fn=None,
                 params_algo=None, **kwargs):
        if g_param is None and sigma_denoiser is not None:
            g_param = sigma_denoiser
        if params_algo is None:
            params_algo = {"lambda": lambda_reg, "stepsize": stepsize, "stepsize_dual": stepsize_dual, "g_param": g_param,
                           "beta": beta, "K": K, "K_adjoint": K_adjoint}
        if trainable_params is None:
            trainable_params = ["lambda", "stepsize", "stepsize_dual", "g_param", "beta"]
        if custom_init is None:
            def custom_init(y, physics):
                x0 = physics.A_adjoint(y)
                return {"est": (x0, x0, y)}
        super().__init__(CPIteration(g_first=g_first, cost_fn=cost_fn), custom_init=custom_init, data_fidelity=data_fidelity,
                         prior=prior, params_algo=params_algo, max_iter=max_iter, crit_conv=crit_conv, thres_conv=thres_conv,
                         early_stop=early_stop, custom_metrics=custom_metrics, unfold=unfold, trainable_params=trainable_params,
                         **kwargs)


_ITERATIONS = {"PGD": PGDIteration, "FISTA": FISTAIteration, "ADMM": ADMMIteration, "HQS": HQSIteration,
               "DRS": DRSIteration, "GD": GDIteration, "CP": CPIteration, "PDCP": CPIteration}


def create_iterator(iteration, prior=None, cost_fn=None, g_first: bool = False, bregman_potential=None, **kwargs):
    """name or instance -> OptimIterator (optimizers.py:884-971); an explicit prior switches the cost on"""
    if prior is None:
        prior = ZeroPrior()
    explicit = prior[0].explicit_prior if isinstance(prior, (list, nn.ModuleList)) else prior.explicit_prior
    has_cost = cost_fn is None and explicit
    if has_cost:
        cost_fn = objective_function
    if isinstance(iteration, str):
        if iteration not in _ITERATIONS:
            raise NotImplementedError(f"iteration {iteration!r} is outside the accelerated path "
                                      f"(have {sorted(_ITERATIONS)})")
        return _ITERATIONS[iteration](g_first=g_first, cost_fn=cost_fn, has_cost=has_cost)
    return iteration


def optim_builder(iteration, max_iter=100, params_algo=None, data_fidelity=None, prior=None, g_first=False, **kwargs):
    """legacy builder (optimizers.py:2560-2679): iteration given by name or as an OptimIterator"""
    iteration = create_iterator(iteration, prior=prior, g_first=g_first)
    return BaseOptim(iteration, max_iter=max_iter, params_algo=params_algo, data_fidelity=data_fidelity, prior=prior,
                     **kwargs)

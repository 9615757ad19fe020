"""Unfolded optimisation (deepinv/unfolded/unfolded.py:9-226): BaseOptim whose stepsize / lambda / g_param are
nn.Parameters and whose forward runs WITHOUT torch.no_grad.  The operator kernels are differentiable (their
backward is the adjoint kernel); the denoiser kernels are inference-only in this round, so an unfolded model
with a libdinvk denoiser is evaluated under torch.no_grad()."""
from __future__ import annotations

from .optim.data_fidelity import L2
from .optim.optim_iterators import OptimIterator
from .optim.optimizers import BaseOptim, DEQConfig, create_iterator


class BaseUnfold(BaseOptim):
    def __init__(self, iterator, params_algo=None, data_fidelity=None, prior=None, max_iter: int = 5,
                 trainable_params=("lambda", "stepsize"), device=None, **kwargs):
        super().__init__(iterator, params_algo=params_algo, data_fidelity=data_fidelity, prior=prior, max_iter=max_iter,
                         unfold=True, trainable_params=list(trainable_params), **kwargs)
        if device is not None:
            self.to(device)


def unfolded_builder(iteration, params_algo=None, data_fidelity=None, prior=None, max_iter: int = 5,
                     trainable_params=("lambda", "stepsize"), device=None, g_first: bool = False, **kwargs):
    if not isinstance(iteration, (str, OptimIterator)):
        raise ValueError("iteration must be a name or an OptimIterator")
    iteration = create_iterator(iteration, prior=prior, g_first=g_first)
    return BaseUnfold(iteration, params_algo=params_algo, data_fidelity=data_fidelity, prior=prior, max_iter=max_iter,
                      trainable_params=trainable_params, device=device, **kwargs)


class BaseDEQ(BaseUnfold):
    """Deep-equilibrium model (deepinv/unfolded/deep_equilibrium.py:11-139): forward = the fixed-point loop without
    gradient tracking, then ONE tracked iteration at the equilibrium; backward = fixed-point sweeps on
    v = J^T v + u (see BaseOptim.DEQ_additional_step).  Same keyword names as the reference's (deprecated) class."""

    def __init__(self, *args, max_iter_backward=50, anderson_acceleration_backward=False, history_size_backward=5,
                 beta_anderson_acc_backward=1.0, eps_anderson_acc_backward=1e-4, jacobian_free=False, **kwargs):
        cfg = DEQConfig(max_iter_backward=max_iter_backward, anderson_acceleration_backward=anderson_acceleration_backward,
                        history_size_backward=history_size_backward, beta_backward=beta_anderson_acc_backward,
                        eps_backward=eps_anderson_acc_backward, jacobian_free=jacobian_free)
        super().__init__(*args, DEQ=cfg, **kwargs)


def DEQ_builder(iteration, params_algo=None, data_fidelity=None, prior=None, cost_fn=None, g_first: bool = False,
                **kwargs):
    """deep_equilibrium.py:142-222"""
    if params_algo is None:
        params_algo = {"lambda": 1.0, "stepsize": 1.0, "g_param": 0.03}
    if data_fidelity is None:
        data_fidelity = L2()
    iterator = create_iterator(iteration, prior=prior, cost_fn=cost_fn, g_first=g_first)
    return BaseDEQ(iterator, data_fidelity=data_fidelity, prior=prior, params_algo=params_algo, **kwargs)

"""Unfolded optimisation (deepinv/unfolded/unfolded.py:9-226): BaseOptim whose stepsize / lambda / g_param are
nn.Parameters and whose forward runs WITHOUT torch.no_grad.  The operator kernels are differentiable (their
backward is the adjoint kernel); the denoiser kernels are inference-only in this round, so an unfolded model
with a libdinvk denoiser is evaluated under torch.no_grad()."""
from __future__ import annotations

from .optim.optim_iterators import ADMMIteration, FISTAIteration, HQSIteration, OptimIterator, PGDIteration
from .optim.optimizers import BaseOptim


class BaseUnfold(BaseOptim):
    def __init__(self, iterator, params_algo=None, data_fidelity=None, prior=None, max_iter: int = 5,
                 trainable_params=("lambda", "stepsize"), device=None, **kwargs):
        super().__init__(iterator, params_algo=params_algo, data_fidelity=data_fidelity, prior=prior, max_iter=max_iter,
                         unfold=True, trainable_params=list(trainable_params), **kwargs)
        if device is not None:
            self.to(device)


def unfolded_builder(iteration, params_algo=None, data_fidelity=None, prior=None, max_iter: int = 5,
                     trainable_params=("lambda", "stepsize"), device=None, g_first: bool = False, **kwargs):
    table = {"PGD": PGDIteration, "FISTA": FISTAIteration, "ADMM": ADMMIteration, "HQS": HQSIteration}
    if isinstance(iteration, str):
        if iteration not in table:
            raise NotImplementedError(f"iteration {iteration!r} is outside the accelerated path")
        iteration = table[iteration](g_first=g_first)
    elif not isinstance(iteration, OptimIterator):
        raise ValueError("iteration must be a name or an OptimIterator")
    return BaseUnfold(iteration, params_algo=params_algo, data_fidelity=data_fidelity, prior=prior, max_iter=max_iter,
                      trainable_params=trainable_params, device=device, **kwargs)

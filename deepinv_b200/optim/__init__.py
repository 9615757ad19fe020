from .data_fidelity import L2, DataFidelity, ZeroFidelity  # noqa: F401
from .graphed import GraphedIteration, HostStreamedIteration  # noqa: F401
from .linear import conjugate_gradient, least_squares  # noqa: F401
from .optim_iterators import (ADMMIteration, FISTAIteration, HQSIteration, OptimIterator,  # noqa: F401
                              PGDIteration)
from .optimizers import ADMM, FISTA, HQS, PGD, BaseOptim, optim_builder  # noqa: F401
from .prior import PnP, Prior, ZeroPrior  # noqa: F401

"""Combining operators: composition A = A_k ... A_1 (`*`, `compose`) and stacking A = [A_1; ...; A_n] (`stack`)
(deepinv/physics/forward.py:73-107, 573-601, 865-987, 1365-1526; deepinv/utils/tensorlist.py).

Pure host-side plumbing: every `A` / `A_adjoint` below is a sequence of the member operators' kernel launches; `prox_l2`
and `A_dagger` of a combined linear operator are the CG on those kernels (LinearPhysics), exactly like the reference,
whose composed / stacked operators also lose the closed forms of their members."""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from .forward import DecomposablePhysics, LinearPhysics, Physics


class TensorList:
    """list of tensors with elementwise arithmetic — the measurement type of stacked operators (utils/tensorlist.py)"""

    def __init__(self, x):
        if isinstance(x, TensorList):
            x = x.x
        self.x = list(x) if isinstance(x, (list, tuple)) else [x]

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i]

    def __setitem__(self, i, v):
        self.x[i] = v

    def __iter__(self):
        return iter(self.x)

    def _zip(self, other, fn):
        if isinstance(other, TensorList):
            return TensorList([fn(a, b) for a, b in zip(self.x, other.x)])
        if isinstance(other, (list, tuple)):
            return TensorList([fn(a, b) for a, b in zip(self.x, other)])
        return TensorList([fn(a, other) for a in self.x])

    def __add__(self, o):
        return self._zip(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._zip(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._zip(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._zip(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._zip(o, lambda a, b: a / b)

    def __neg__(self):
        return TensorList([-a for a in self.x])

    def to(self, *args, **kwargs):
        return TensorList([a.to(*args, **kwargs) for a in self.x])

    def flatten(self):
        return torch.cat([a.reshape(a.shape[0], -1) for a in self.x], dim=1)

    @property
    def shape(self):
        return [a.shape for a in self.x]


# ---- composition ------------------------------------------------------------------------------------------------
class ComposedPhysics(Physics):
    r"""y = N_k(A_k(... A_1(x)))  (forward.py:865-933); keeps the noise / sensor model of the last operator"""

    def __init__(self, *physics, **kwargs):
        super().__init__(**{k: v for k, v in kwargs.items() if k in ("max_iter", "tol", "solver")})
        self.physics_list = nn.ModuleList([])
        for p in physics:
            self.physics_list.extend(p.physics_list if isinstance(p, ComposedPhysics) else [p])
        self.noise_model = physics[-1].noise_model
        self.sensor_model = physics[-1].sensor_model

    def A(self, x, **kwargs):
        for p in self.physics_list:
            x = p.A(x, **kwargs)
        return x

    def update_parameters(self, **kwargs):
        for p in self.physics_list:
            p.update_parameters(**kwargs)

    def __getitem__(self, item):
        return self.physics_list[item]

    def __str__(self):
        return "ComposedPhysics(" + "\n".join(f"{p}" for p in reversed(self.physics_list)) + ")"

    __repr__ = __str__


class ComposedLinearPhysics(ComposedPhysics, LinearPhysics):
    r"""A = A_k ... A_1, A^T = A_1^T ... A_k^T (forward.py:936-967)"""

    def __init__(self, *physics, **kwargs):
        ComposedPhysics.__init__(self, *physics, **kwargs)
        self.solver = "CG"

    def A_adjoint(self, y, **kwargs):
        for p in reversed(self.physics_list):
            y = p.A_adjoint(y, **kwargs)
        return y


def compose(*physics, **kwargs):
    """A = physics[-1] o ... o physics[0]  (forward.py:970-987)"""
    if any(isinstance(p, DecomposablePhysics) for p in physics):
        warnings.warn("At least one input physics is a DecomposablePhysics, but resulting physics will not be decomposable. "
                      "`A_dagger` and `prox_l2` will fall back to approximate methods, which may impact performance.")
    if all(isinstance(p, LinearPhysics) for p in physics):
        return ComposedLinearPhysics(*physics, **kwargs)
    return ComposedPhysics(*physics, **kwargs)


# ---- stacking ---------------------------------------------------------------------------------------------------
class StackedPhysics(Physics):
    r"""y = [A_1(x), ..., A_n(x)] as a TensorList (forward.py:1380-1476)"""

    def __init__(self, physics_list, **kwargs):
        super().__init__()
        self.physics_list = nn.ModuleList([])
        for p in physics_list:
            self.physics_list.extend(p.physics_list if isinstance(p, StackedPhysics) else [p])

    def A(self, x, **kwargs):
        return TensorList([p.A(x, **kwargs) for p in self.physics_list])

    def __getitem__(self, item):
        return self.physics_list[item]

    def __len__(self):
        return len(self.physics_list)

    def sensor(self, y, **kwargs):
        for i, p in enumerate(self.physics_list):
            y[i] = p.sensor(y[i], **kwargs)
        return y

    def noise(self, y, **kwargs):
        for i, p in enumerate(self.physics_list):
            y[i] = p.noise(y[i], **kwargs)
        return y

    def set_noise_model(self, noise_model, item=0):
        self.physics_list[item].set_noise_model(noise_model)

    def update_parameters(self, **kwargs):
        for p in self.physics_list:
            p.update_parameters(**kwargs)

    def __str__(self):
        return "StackedPhysics(" + "\n".join(f"{p}" for p in self.physics_list) + ")"

    __repr__ = __str__


class StackedLinearPhysics(StackedPhysics, LinearPhysics):
    r"""A^T y = sum_i A_i^T y_i (forward.py:1479-1526); A^T A = sum_i A_i^T A_i uses each member's fused normal operator"""

    def __init__(self, physics_list, reduction="sum", **kwargs):
        StackedPhysics.__init__(self, physics_list, **kwargs)
        self.solver = "CG"
        if reduction == "sum":
            self.reduction = sum
        elif reduction == "mean":
            self.reduction = lambda x: sum(x) / len(x)
        elif reduction in ("none", None):
            self.reduction = lambda x: x
        else:
            raise ValueError("reduction must be either sum, mean or none.")
        if reduction != "sum":
            warnings.warn(f"Using `reduction={reduction}` is deprecated and breaks the adjointness property of the operator.",
                          DeprecationWarning, stacklevel=2)
        self._is_sum = reduction == "sum"

    def A_adjoint(self, y, **kwargs):
        return self.reduction([p.A_adjoint(y[i], **kwargs) for i, p in enumerate(self.physics_list)])

    def A_adjoint_A(self, x, **kwargs):
        if not self._is_sum:
            return self.A_adjoint(self.A(x, **kwargs), **kwargs)
        return sum(p.A_adjoint_A(x, **kwargs) for p in self.physics_list)


def stack(*physics):
    """[A_1; ...; A_n]  (forward.py:1365-1377)"""
    if all(isinstance(p, LinearPhysics) for p in physics):
        return StackedLinearPhysics(physics)
    return StackedPhysics(physics)

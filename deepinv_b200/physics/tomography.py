"""Parallel-beam Tomography on libdinvk's Radon kernels.

Drop-in for deepinv/physics/tomography.py:26-350 (parallel beam; `fan_beam=True` and the Astra
wrapper are out of scope).  Same constructor keywords, buffers (`angles`,
`operator_norm`), shape errors and output conventions: `A` returns the sinogram as the (B,C,P,A) view
of angle-major memory exactly like the reference (radon.py:291-293), `A_adjoint` is the exact
transpose when `adjoint_via_backprop=True` (the reference obtains it from autograd) and the IRadon
back-projection otherwise, `A_dagger(fbp=True)` is ramp filter + adjoint with the reference's
scalings (:258-293), `A_dagger()` / `prox_l2` are CG on the normal equations (forward.py:751-862).
"""
from __future__ import annotations

import math
from warnings import warn

import numpy as np
import torch

from .. import ops
from .forward import LinearPhysics, TensorKey, cache_hit, linear_apply


def _deg2rad(theta: torch.Tensor) -> torch.Tensor:
    return theta * 4 * torch.ones(1, device=theta.device, dtype=theta.dtype).atan() / 180  # radon.py:70-71


class Tomography(LinearPhysics):
    def __init__(self, angles, img_width: int, circle: bool = False, parallel_computation: bool = True,
                 adjoint_via_backprop: bool = True, fbp_interpolate_boundary: bool = False, normalize: bool | None = None,
                 fan_beam: bool = False, fan_parameters: dict = None, device="cpu", dtype: torch.dtype = torch.float,
                 **kwargs):
        super().__init__(device=device, **kwargs)
        if dtype != torch.float:
            raise NotImplementedError("deepinv_b200.Tomography computes in float32")
        if isinstance(angles, int):
            angles = torch.linspace(0, 180, steps=angles + 1, device=device)[:-1].to(device)
        elif isinstance(angles, (list, tuple, np.ndarray)):
            angles = torch.tensor(angles).to(device)
        elif not isinstance(angles, torch.Tensor):
            raise ValueError(f"angles must be int, float, iterable or Tensor, but got {type(angles)}")
        self.register_buffer("angles", angles)
        self.fan_beam = bool(fan_beam)
        self.fan_parameters = None
        if self.fan_beam:
            # defaults of radon.py:224-240; the grid is built on the padded image of size G = W (circle) or ceil(sqrt(2) W)
            fp = dict(fan_parameters or {})
            fp.setdefault("pixel_spacing", 0.5 / img_width)
            fp.setdefault("source_radius", 57.5)
            fp.setdefault("detector_radius", 57.5)
            fp.setdefault("n_detector_pixels", 258)
            fp.setdefault("detector_spacing", 0.077)
            self.fan_parameters = fp
        self.adjoint_via_backprop = adjoint_via_backprop
        if circle and fbp_interpolate_boundary:
            warn("The argument fbp_interpolate_boundary=True is not applicable if circle=True. The value "
                 "fbp_interpolate_boundary will be changed to False...")
            fbp_interpolate_boundary = False
        self.fbp_interpolate_boundary = fbp_interpolate_boundary
        self.img_width = img_width
        self.circle = circle
        self.dtype = dtype
        # P = ceil(sqrt(2) W) in float32 (radon.py:60-61, 319)
        self.P = img_width if circle else int(((2 * torch.ones(1)).sqrt() * img_width).ceil())
        self.G = self.P  # grid size of the padded image
        if self.fan_beam:
            fp = self.fan_parameters
            self.P = int(fp["n_detector_pixels"])  # detector cells of a sinogram row
            sf = 2.0 / (self.G * fp["pixel_spacing"])  # radon.py:23-28, python floats like the reference
            src, det, spacing = fp["source_radius"] * sf, fp["detector_radius"] * sf, fp["detector_spacing"] * sf
            self._fan = (0.5 * (spacing * (self.P - 1)), src, src + det)
        self._trig_key = None
        if normalize is None:
            warn("The default value of `normalize` is not specified and will be automatically set to `True`. Set "
                 "`normalize` explicitly to `True` or `False` to avoid this warning.")
            normalize = True
        self.normalize = False
        if normalize:
            x0 = torch.randn((1, img_width, img_width), generator=torch.Generator(self.angles.device).manual_seed(0),
                             device=self.angles.device)[None]
            operator_norm = self.compute_sqnorm(x0, verbose=False).sqrt()
            self.register_buffer("operator_norm", operator_norm)
            self.normalize = True
        self.to(device)

    # ---- geometry tables -------------------------------------------------------------------------------
    def _trig(self):
        a = self.angles
        if not cache_hit(self._trig_key, a):
            # cos / sin of the (fp32) angles evaluated in fp64 and handed to the kernels as fp32 (hi, lo) pairs, shape (2, A):
            # the parallel-beam kernels compute sample coordinates in fp64 (csrc/radon.cu header); row 0 alone is the fp32 table
            th = _deg2rad(a.to(torch.float32).to(torch.float64))
            split = lambda v: torch.stack([v.float(), (v - v.float().double()).float()]).contiguous()
            self._cos, self._sin = split(th.cos()), split(th.sin())
            self._trig_key = TensorKey(a)
        return self._cos, self._sin

    def _norm(self) -> float:
        return float(self.operator_norm) if self.normalize else 1.0

    @property
    def theta(self):
        warn("The attribute `theta` is deprecated and will be removed in a future version. Use `angles` instead.",
             DeprecationWarning, stacklevel=2)
        return self.angles

    # ---- raw kernels (angle-major sinograms) ---------------------------------------------------------------
    def _fwd_am(self, x, scale):
        c, s = self._trig()
        if self.fan_beam:
            return ops.fanbeam(x, self.img_width, self.G, self.P, c, s, self.circle, *self._fan, scale, adjoint=False)
        return ops.radon_fwd(x, self.P, c, s, self.circle, scale)

    def _adj_am(self, y_am, scale, iradon=False):
        c, s = self._trig()
        if self.fan_beam:  # always the exact transpose (tomography.py:322-342: fan_beam forces the autograd adjoint)
            return ops.fanbeam(y_am, self.img_width, self.G, self.P, c, s, self.circle, *self._fan, scale, adjoint=True)
        return ops.radon_adj(y_am, self.img_width, c, s, self.circle, scale, iradon=iradon)

    @staticmethod
    def _to_am(y):
        """(B,C,P,A) in any layout -> contiguous angle-major (B,C,A,P); free when y is the view A() returned"""
        return y.transpose(-2, -1).contiguous()

    def _A(self, x):
        return self._fwd_am(x, 1.0 / self._norm()).transpose(-2, -1)

    def _At(self, y):
        n = self._norm()
        if self.adjoint_via_backprop or self.fan_beam:
            return self._adj_am(self._to_am(y), 1.0 / n)
        # ApplyRadon adjoint = iradon(y, filtering=False) / pi * 2A (radon.py:512-514), iradon itself carries pi/(2A)
        return self._adj_am(self._to_am(y), 1.0 / n, iradon=True)

    def A(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        if not tuple(x.shape[-2:]) == (self.img_width, self.img_width):
            raise ValueError(f"Input image size {tuple(x.shape[-2:])} does not match the operator image size "
                             f"{(self.img_width, self.img_width)}.")
        return linear_apply(x, self._A, self._At)

    def A_adjoint(self, y: torch.Tensor, **kwargs) -> torch.Tensor:
        if y.dim() != 4 or y.shape[-2] != self.P or y.shape[-1] != self.angles.numel():
            raise ValueError(f"expected a sinogram of shape (B,C,{self.P},{self.angles.numel()}), got {tuple(y.shape)}")
        return linear_apply(y, self._At, self._A)

    def fbp(self, y: torch.Tensor, **kwargs) -> torch.Tensor:
        """filtered back-projection (tomography.py:258-293)"""
        A = self.angles.numel()
        yf = ops.ramp_filter(self._to_am(y))
        n = self._norm()
        if self.adjoint_via_backprop or self.fan_beam:
            # A_adjoint(filter(y)) * pi/(2A) * norm^2  with A_adjoint = R^T / norm
            out = self._adj_am(yf, (math.pi / (2 * A)) * n)
        else:
            # IRadon (incl. its pi/(2A)) * norm
            out = self._adj_am(yf, (math.pi / (2 * A)) * n, iradon=True)
        if self.fbp_interpolate_boundary:
            out = torch.nn.functional.pad(out[:, :, 2:-2, 2:-2], (2, 2, 2, 2), mode="replicate")
        return out

    def A_dagger(self, y: torch.Tensor, fbp: bool = False, **kwargs) -> torch.Tensor:
        if fbp:
            return self.fbp(y, **kwargs)
        return super().A_dagger(y, **kwargs)

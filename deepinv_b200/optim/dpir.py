"""DPIR: the library's flagship plug-and-play recipe (deepinv/optim/dpir.py:11-81) — 8 HQS iterations with a PnP
DRUNet prior, the denoiser level decaying log-uniformly from 49/255 to the noise level of the measurement and the
stepsize tied to it.  Each iteration is `prox_l2` (closed-form spectral kernel for MRI / BlurFFT, CG on the
operator kernels otherwise) followed by one denoiser pass on the tensor-core (or fp32) convolution kernels."""
from __future__ import annotations

import torch

from .data_fidelity import L2
from .optimizers import BaseOptim, create_iterator
from .prior import PnP


def get_DPIR_params(noise_level_img: float, device="cpu"):
    """(sigma_denoiser per iteration, stepsize per iteration, max_iter)  (dpir.py:11-35); the schedule is computed on
    the host in fp32 exactly like the reference (torch.logspace on CPU) and then moved"""
    max_iter = 8
    s1, s2 = 49.0 / 255.0, float(noise_level_img)
    sigma_denoiser = torch.logspace(torch.log10(torch.tensor(s1, dtype=torch.float32)),
                                    torch.log10(torch.tensor(s2, dtype=torch.float32)), steps=max_iter,
                                    dtype=torch.float32, device="cpu").to(device)
    stepsize = (sigma_denoiser / max(0.01, s2)) ** 2
    lamb = 1 / 0.23
    return sigma_denoiser, lamb * stepsize, max_iter


class DPIR(BaseOptim):
    def __init__(self, sigma=0.1, denoiser=None, device="cpu"):
        if denoiser is None:
            raise RuntimeError("deepinv_b200.DPIR: pass a denoiser (the reference downloads pretrained DRUNet weights; "
                               "there is no network here) — e.g. DRUNet(pretrained=<checkpoint path>)")
        prior = PnP(denoiser=denoiser)
        sigma_denoiser, stepsize, max_iter = get_DPIR_params(float(sigma), device=device)
        super().__init__(create_iterator("HQS", prior=prior, cost_fn=None, g_first=False), max_iter=max_iter,
                         data_fidelity=L2(), prior=prior, early_stop=False,
                         params_algo={"stepsize": stepsize, "g_param": sigma_denoiser})

"""DnCNN on libdinvk's convolution kernels (drop-in for deepinv/models/dncnn.py:14-144).

Same module tree (`in_conv`, `conv_list`, `out_conv`) so reference state_dicts load unchanged;
forward = out_conv(relu(conv(...relu(in_conv(x))))) + x (:121-138) with bias + ReLU fused into each
convolution launch and the residual fused into the last one.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .base import Denoiser, _no_grad_guard


def weights_init_kaiming(m):
    if m.__class__.__name__.find("Conv") != -1:
        nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")


class DnCNN(Denoiser):
    def __init__(self, in_channels: int = 3, out_channels: int = 3, depth: int = 20, bias: bool = True, nf: int = 64,
                 pretrained: str | None = "download", pretrained_2d_isotropic: bool = False, device="cpu", dim=2,
                 precision: str = "fp32"):
        super().__init__()
        if int(dim) != 2:
            raise NotImplementedError("deepinv_b200.DnCNN is 2-D only")
        self.depth = depth
        self.in_conv = nn.Conv2d(in_channels, nf, 3, 1, 1, bias=bias)
        self.conv_list = nn.ModuleList([nn.Conv2d(nf, nf, 3, 1, 1, bias=bias) for _ in range(depth - 2)])
        self.out_conv = nn.Conv2d(nf, out_channels, 3, 1, 1, bias=bias)
        self.nl_list = nn.ModuleList([nn.ReLU() for _ in range(depth - 1)])
        self.precision = precision
        self._tc = None
        self._tc32 = None
        if pretrained is not None:
            if pretrained.startswith("download"):
                raise RuntimeError("pretrained weights cannot be downloaded here (no network): pass a checkpoint path "
                                   "or pretrained=None and load a state_dict")
            self.load_state_dict(torch.load(pretrained, map_location="cpu"), strict=True)
            self.eval()
        else:
            self.apply(weights_init_kaiming)
        if device is not None:
            self.to(device)

    def forward(self, x: torch.Tensor, sigma=None) -> torch.Tensor:
        if self.precision == "bf16":
            _no_grad_guard("DnCNN(precision='bf16')", x, self.in_conv.weight)
            from .tc_engine import dncnn_forward_bf16

            return dncnn_forward_bf16(self, x)
        if self.precision in ("tc32", "tc32h"):
            _no_grad_guard(f"DnCNN(precision='{self.precision}')", x, self.in_conv.weight)
            from .tc_engine import dncnn_forward_tc32

            return dncnn_forward_tc32(self, x)
        t = ops.conv_f32_ag(x, self.in_conv.weight, bias=self.in_conv.bias, relu=True)
        for conv in self.conv_list:
            t = ops.conv_f32_ag(t, conv.weight, bias=conv.bias, relu=True)
        return ops.conv_f32_ag(t, self.out_conv.weight, bias=self.out_conv.bias, res=x)

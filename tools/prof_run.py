#!/usr/bin/env python
"""Small workloads for ncu captures: `drunet` (one bf16 DRUNet forward, cfg2 shape) or `mri` (A, A^T, data step)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import deepinv_b200 as dinv  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "drunet"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
with torch.no_grad():
    if what == "drunet":
        den = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None, precision="bf16").to(dev).eval()
        x = torch.randn(64, 2, 256, 256, device=dev)
        for _ in range(reps):
            den(x, 0.05)
    else:
        B, H, W = 64, 256, 256
        x = torch.randn(B, 2, H, W, device=dev)
        cols = (torch.rand(B, 1, 1, W, device=dev) > 0.75).float().expand(B, 2, H, W).contiguous()
        p = dinv.physics.MRI(mask=cols, img_size=(2, H, W), device=dev)
        for _ in range(reps):
            y = p.A(x)
            aty = p.A_adjoint(y)
            p.normal_step(x, aty, 1.0)
    torch.cuda.synchronize()

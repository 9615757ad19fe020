#!/usr/bin/env python
"""One or two launches of each operator kernel family at the BASELINE.json sizes — the command ncu wraps for profiles/."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import deepinv_b200 as dinv  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1:] or ["mri", "tomo", "blur", "mcmri"]
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    if "mri" in which:
        B, H, W = 64, 256, 256
        x = torch.randn(B, 2, H, W, device=dev, generator=g)
        cols = (torch.rand(B, 1, 1, W, device=dev, generator=g) > 0.75).float().expand(B, 2, H, W).contiguous()
        p = dinv.physics.MRI(mask=cols, img_size=(2, H, W), device=dev)
        for _ in range(2):
            y = p.A(x)
            p.A_adjoint(y)
            p.normal_step(x, y, 1.0)
    if "tomo" in which:
        p = dinv.physics.Tomography(angles=180, img_width=512, normalize=False, device=dev)
        x = torch.randn(32, 1, 512, 512, device=dev, generator=g)
        for _ in range(2):
            y = p.A(x)
            p.A_adjoint(y)
    if "blur" in which:
        x = torch.rand(32, 1, 1024, 1024, device=dev, generator=g)
        f = torch.rand(1, 1, 31, 31, device=dev, generator=g)
        f /= f.sum()
        p = dinv.physics.Blur(filter=f, padding="circular", device=dev)
        pf = dinv.physics.BlurFFT(img_size=(1, 1024, 1024), filter=f, device=dev)
        for _ in range(2):
            p.A(x)
            pf.A(x)
    if "mcmri" in which:
        B, N, H, W = 32, 8, 320, 320
        x = torch.randn(B, 2, H, W, device=dev, generator=g)
        maps = torch.view_as_complex(torch.randn(1, N, H, W, 2, device=dev, generator=g))
        maps = maps / maps.abs().pow(2).sum(1, keepdim=True).sqrt()
        mask = (torch.rand(B, 1, 1, W, device=dev, generator=g) > 0.875).float().expand(B, 2, H, W).contiguous()
        p = dinv.physics.MultiCoilMRI(mask=mask, coil_maps=maps, img_size=(2, H, W), device=dev)
        for _ in range(2):
            y = p.A(x)
            p.A_adjoint(y)
torch.cuda.synchronize()

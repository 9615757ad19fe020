#!/usr/bin/env python
"""Launch the tc32 halo-reuse convolution a few times at the bench shape (64 -> 64, 64 x 256 x 256) — the command ncu wraps."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deepinv_b200 import ops  # noqa: E402
from deepinv_b200.models.tc_engine import _pack3x3_slab_tc32  # noqa: E402

dev = torch.device("cuda:0")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fmt = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H = {64: 256, 128: 128, 256: 64, 512: 32}[C]
x = ops.nchw_to_split16(torch.randn(64, C, H, H, device=dev).abs_(), fmt)
r = ops.nchw_to_split16(torch.randn(64, C, H, H, device=dev), fmt)
w = _pack3x3_slab_tc32(torch.randn(C, C, 3, 3, device=dev) / (3 * C ** 0.5), fmt)
for _ in range(3):
    ops.conv_tc32_slab(x, w, C, relu=True)
    ops.conv_tc32_slab(x, w, C, res=r)
torch.cuda.synchronize()

#!/usr/bin/env python
"""Time single DRUNet-shaped 3x3 layers on the tensor-core path (CUDA events, 20 reps each)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deepinv_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, C, H) in [(64, 64, 256), (64, 128, 128), (64, 256, 64), (64, 512, 32)]:
    x = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
    r = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(C, 9 * C, device=dev) / (3 * C ** 0.5)).to(torch.bfloat16)
    for tag, kw in (("relu", dict(relu=True)), ("res", dict(res=r))):
        for _ in range(3):
            ops.conv3x3_bf16(x, w, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv3x3_bf16(x, w, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gf = 2 * B * H * H * C * 9 * C / 1e9
        print(f"C={C:4d} H={H:4d} {tag:5s} {ms * 1e3:8.1f} us  {gf / ms:8.1f} TFLOP/s", flush=True)


def _time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# head (3 -> 64 from NCHW fp32) and tail (64 -> 2 to NCHW fp32) of DRUNet at the bench shape
B, H = 64, 256
x0 = torch.randn(B, 3, H, H, device=dev)
w64 = (torch.randn(64, 64, device=dev) / 5).to(torch.bfloat16)
ms = _time(lambda: ops.conv3x3_head_bf16(x0, w64))
print(f"head 3->64 H={H}: {ms * 1e3:8.1f} us  ({(x0.numel() * 4 + B * H * H * 64 * 2) / ms / 1e6:7.1f} GB/s algorithmic)", flush=True)
t = torch.randn(B, H, H, 64, device=dev).to(torch.bfloat16)
w16 = (torch.randn(16, 9 * 64, device=dev) / 24).to(torch.bfloat16)
ms = _time(lambda: ops.conv3x3_bf16_tail(t, w16, 2))
print(f"tail 64->2 H={H}: {ms * 1e3:8.1f} us  ({(t.numel() * 2 + B * 2 * H * H * 4) / ms / 1e6:7.1f} GB/s algorithmic)", flush=True)

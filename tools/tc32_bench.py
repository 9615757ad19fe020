#!/usr/bin/env python
"""Time the fp32-grade tensor-core path (precision="tc32") at the DRUNet shapes of the bench (batch 64, 256x256):
single layers (CUDA events, 10 reps, operands > L2) and whole forwards of the three precisions.
    python tools/tc32_bench.py [layers] [net] [window=W]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import deepinv_b200 as dinv  # noqa: E402
from deepinv_b200 import ops  # noqa: E402
from deepinv_b200.models.tc_engine import _pack3x3_slab_tc32, _pack3x3_tc32, _pack_down_tc32, _pack_up_tc32  # noqa: E402

dev = torch.device("cuda:0")
args = sys.argv[1:] or ["layers", "net"]
windows = [int(a.split("=")[1]) for a in args if a.startswith("window=")] or [0]
fmts = [int(a.split("=")[1]) for a in args if a.startswith("fmt=")] or [0, 1]


def _time(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def split_randn(B, H, W, C, fmt=0):
    return ops.nchw_to_split16(torch.randn(B, C, H, W, device=dev).abs_(), fmt)


torch.manual_seed(0)
B = 64
if "layers" in args:
    for fmt, (C, H) in [(f, ch) for f in fmts for ch in [(64, 256), (128, 128), (256, 64), (512, 32)]]:
        x = split_randn(B, H, H, C, fmt)
        r = split_randn(B, H, H, C, fmt)
        w4 = torch.randn(C, C, 3, 3, device=dev) / (3 * C ** 0.5)
        w = _pack3x3_tc32(w4, fmt)
        ws = _pack3x3_slab_tc32(w4, fmt)
        gf = 2 * B * H * H * C * 9 * C / 1e9
        for win in (windows if "slab" in args else []):
            for tag, kw in (("relu", dict(relu=True)), ("res", dict(res=r))):
                ms = _time(lambda: ops.conv_tc32_slab(x, ws, C, window=win, **kw))
                print(json.dumps({"op": f"conv3x3 tc32 SLAB fmt={fmt} {C}->{C} @{H}x{H} B={B} {tag}", "window_blocks": win, "us": ms * 1e3,
                                  "TFLOPs_fp32_equiv": gf / ms}), flush=True)
        for win in (windows if "tap" in args else []):
            for tag, kw in (("relu", dict(relu=True)), ("res", dict(res=r))):
                ms = _time(lambda: ops.conv_tc32(x, w, C, window=win, **kw))
                print(json.dumps({"op": f"conv3x3 tc32 fmt={fmt} {C}->{C} @{H}x{H} B={B} {tag}", "window": win, "us": ms * 1e3,
                                  "TFLOPs_fp32_equiv": gf / ms}), flush=True)
        del x, r, w
    for fmt, (Ci, Co, H) in [(f, ch) for f in fmts for ch in [(64, 128, 256), (128, 256, 128), (256, 512, 64)]]:
        x = split_randn(B, H, H, Ci, fmt)
        wd = _pack_down_tc32(torch.randn(Co, Ci, 2, 2, device=dev) / (2 * Ci ** 0.5), fmt)
        ms = _time(lambda: ops.conv_tc32(x, wd, Co, kind=1))
        gf = 2 * B * (H // 2) ** 2 * 4 * Ci * Co / 1e9
        print(json.dumps({"op": f"down 2x2 fmt={fmt} {Ci}->{Co} @{H}", "us": ms * 1e3, "TFLOPs_fp32_equiv": gf / ms}), flush=True)
        xs = split_randn(B, H // 2, H // 2, Co, fmt)
        wu = _pack_up_tc32(torch.randn(Co, Ci, 2, 2, device=dev) / (Co ** 0.5), fmt)
        ms = _time(lambda: ops.conv_tc32(xs, wu, Ci, kind=2))
        print(json.dumps({"op": f"up 2x2 fmt={fmt} {Co}->{Ci} @{H // 2}", "us": ms * 1e3, "TFLOPs_fp32_equiv": gf / ms}), flush=True)
        del x, xs
    for fmt in fmts:
        x0 = torch.randn(B, 3, 256, 256, device=dev)
        wh = torch.randn(64, 3, 3, 3, device=dev) / 5
        ms = _time(lambda: ops.conv_tc32_head(x0, wh, fmt=fmt))
        eb = 4 if fmt else 8
        print(json.dumps({"op": f"head fmt={fmt} 3->64 @256", "us": ms * 1e3, "GBps": (x0.numel() * 4 + B * 256 * 256 * 64 * eb) / ms / 1e6}), flush=True)
        t = split_randn(B, 256, 256, 64, fmt)
        wt = torch.randn(2, 64, 3, 3, device=dev) / 24
        ms = _time(lambda: ops.conv_tc32_tail(t, wt))
        print(json.dumps({"op": f"tail fmt={fmt} 64->2 @256", "us": ms * 1e3, "GBps": (B * 256 * 256 * 64 * eb + B * 2 * 256 * 256 * 4) / ms / 1e6}), flush=True)
        del t, x0

if "net" in args:
    x = torch.randn(B, 2, 256, 256, device=dev)
    for prec in ("tc32", "tc32h", "bf16"):
        torch.manual_seed(0)
        m = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None, precision=prec).to(dev).eval()
        with torch.no_grad():
            ms = _time(lambda: m(x, 0.05), reps=5)
        print(json.dumps({"op": f"DRUNet forward B={B} 256x256 precision={prec}", "ms": ms, "TFLOPs_fp32_equiv": 277.4 * B / ms}), flush=True)
    # reconstruction-level error of the precisions against the fp32 CUDA-core path on 4 images
    with torch.no_grad():
        xs = x[:4]
        outs = {}
        for prec in ("fp32", "tc32", "tc32h", "bf16"):
            torch.manual_seed(0)
            m = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None, precision=prec).to(dev).eval()
            outs[prec] = m(xs, 0.05)
        for prec in ("tc32", "tc32h", "bf16"):
            e = float((outs[prec] - outs["fp32"]).norm() / outs["fp32"].norm())
            print(json.dumps({"op": f"DRUNet rel. L2 error {prec} vs fp32 path (4 x 256x256)", "err": e}), flush=True)

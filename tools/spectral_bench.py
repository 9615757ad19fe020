#!/usr/bin/env python
"""MRI operator timings at cfg2 (64 x 256^2) with inputs LARGER than L2: each operator is captured into one CUDA graph
that walks `SETS` disjoint operand sets (SETS x >= 67 MB > 126 MB L2), so every call reads its operands from HBM.
`--hot` re-uses one set (L2-resident upper bound).  One JSON line per operator (algorithmic bytes / time vs the measured
HBM peak).  `--eager N` runs N plain calls of each operator instead (for ncu captures)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import deepinv_b200 as dinv  # noqa: E402

dev = torch.device("cuda:0")
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
HOT = "--hot" in sys.argv
EAGER = int(sys.argv[sys.argv.index("--eager") + 1]) if "--eager" in sys.argv else 0
SETS = 1 if HOT else 6
B, H, W = 64, 256, 256
REPLAYS = 20


def graph_time(calls):
    """calls: list of thunks (one per operand set).  Returns ms per call."""
    for c in calls:
        c()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for c in calls:
            c()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for c in calls:
            c()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPLAYS):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (REPLAYS * len(calls))


def main():
    gen = torch.Generator(device=dev).manual_seed(0)
    img_mb = B * 2 * H * W * 4 / 1e6
    with torch.no_grad():
        for tag, full in (("line mask", False), ("full fp32 mask", True)):
            sets = []
            for _ in range(SETS):
                x = torch.randn(B, 2, H, W, device=dev, generator=gen)
                if full:
                    m = (torch.rand(B, 2, H, W, device=dev, generator=gen) > 0.75).float()
                else:
                    m = (torch.rand(B, 1, 1, W, device=dev, generator=gen) > 0.75).float().expand(B, 2, H, W).contiguous()
                p = dinv.physics.MRI(mask=m, img_size=(2, H, W), device=dev)
                y = p.A(x)
                aty = p.A_adjoint(y)
                sets.append((p, x, y, aty))
            extra = img_mb if full else 0.0
            ops_ = [
                ("MRI.A", lambda p, x, y, aty: p.A(x), 2 * img_mb + extra),
                ("MRI.A_adjoint", lambda p, x, y, aty: p.A_adjoint(y), 2 * img_mb + extra),
                ("MRI.A_adjoint_A", lambda p, x, y, aty: p.A_adjoint_A(x), 2 * img_mb + extra),
                ("MRI.normal_step", lambda p, x, y, aty: p.normal_step(x, aty, 1.0), 3 * img_mb + extra),
                ("MRI.prox_l2", lambda p, x, y, aty: p.prox_l2(x, y, 1.0), 3 * img_mb + extra),
            ]
            for name, fn, mb in ops_:
                if EAGER:
                    for _ in range(EAGER):
                        for st in sets[:1]:
                            fn(*st)
                    torch.cuda.synchronize()
                    continue
                ms = graph_time([(lambda st=st: fn(*st)) for st in sets])
                print(json.dumps({"op": f"{name} 64x256^2 [{tag}]", "ms": round(ms, 5), "algorithmic_MB": round(mb, 1),
                                  "GBps": round(mb / ms, 1), "frac_hbm": round(mb / ms / PEAK, 4),
                                  "operands": "L2-resident (1 set)" if HOT else f"{SETS} disjoint sets per graph (> L2)"}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — PnP-PGD iterations/s on MRI 256x256, 4x Cartesian mask, DRUNet denoiser (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # the B200 arm (default)
    python bench.py --impl reference --steps 2 --warmup 1          # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU, batch sharded (weak scaling)

One "step" = one PnP-PGD iteration over the whole batch: fused L2 data step
z = x - gamma (A^T A x - A^T y) followed by x = DRUNet(z, sigma), through the package's public
optimiser API (deepinv_b200.optim.PGD.single_iteration).  Prints ONE JSON line (contract in the task
statement): `value` with inputs resident in HBM, `e2e` through host buffers, `roofline` for the
dominant kernel family (the denoiser convolutions) plus per-operator HBM fractions under
`operators`, `cpu_baseline` (oracle timed on the host cores), clocks sampled during the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H = W = 256
BATCH = 64
ACCEL = 4
SIGMA_DEN = 0.05
STEPSIZE = 1.0
DRUNET_GFLOP_PER_IMAGE = 277.40  # SURVEY Appendix A.12 (C=2, 256x256)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def cartesian_mask(batch: int, h: int, w: int, accel: int, seed: int) -> torch.Tensor:
    """random Cartesian line masks like RandomMaskGenerator (generator/mri.py:136-196): a fully sampled centre
    band (8 % for 4x, 4 % for 8x) plus uniformly random columns up to 1/accel density, constant along H"""
    g = torch.Generator().manual_seed(seed)
    center = {4: 0.08, 8: 0.04}.get(accel, 0.32 / accel)
    n_center = int(round(w * center))
    n_total = int(round(w / accel))
    m = torch.zeros(batch, 1, 1, w)
    lo = (w - n_center) // 2
    for b in range(batch):
        m[b, 0, 0, lo: lo + n_center] = 1
        rest = torch.tensor([i for i in range(w) if not (lo <= i < lo + n_center)])
        pick = rest[torch.randperm(len(rest), generator=g)[: n_total - n_center]]
        m[b, 0, 0, pick] = 1
    return m.expand(batch, 2, h, w).contiguous()


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 20` running from before the warm-up; every sample carries the driver's timestamp, so the samples that
    fall INSIDE a host-side window (the timed region, or — when that is shorter than a few sampler periods — a replay of the
    same workload right after it) can be picked out afterwards."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    @staticmethod
    def now():
        import datetime

        return datetime.datetime.now()

    def stop(self, windows) -> dict:
        """windows: list of (label, t_begin, t_end) in preference order; the first one holding >= 3 samples is reported"""
        import datetime

        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = []
        for r in Path(self.f.name).read_text().strip().splitlines():
            c = [v.strip() for v in r.split(",")]
            if len(c) < 9:
                continue
            try:
                rows.append((datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f"), c))
            except ValueError:
                continue
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        pick, label = rows, "whole run (no sample inside the windows)"
        for lab, t0, t1 in windows:
            inside = [rc for rc in rows if t0 <= rc[0] <= t1]
            if len(inside) >= 3:
                pick, label = inside, lab
                break
        cols = [c for _, c in pick]
        sm = sorted(float(c[1]) for c in cols)
        reasons = set()
        for c in cols:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if c[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(cols[0][2]), "power_w_max": max(float(c[3]) for c in cols),
                "samples": len(cols), "window": label, "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's CPU path
# ---------------------------------------------------------------------------------------------
def cpu_pgd_iteration_seconds(sample_batch: int, repeats: int, threads: int):
    """time ONE PnP-PGD iteration of the oracle (torch-CPU restatement of the reference path) on
    `sample_batch` images; returns best-of-`repeats` seconds"""
    from oracle import ref_ops as R

    torch.set_num_threads(threads)
    torch.manual_seed(0)
    import deepinv_b200 as dinv

    den = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None)  # parameter container only (random init, seed 0)
    sd = {k: v.detach() for k, v in den.state_dict().items()}
    x = torch.randn(sample_batch, 2, H, W)
    mask = cartesian_mask(sample_batch, H, W, ACCEL, seed=0)
    y = R.mri_A(x, mask)
    best = float("inf")
    with torch.no_grad():
        xk = R.mri_At(y, mask)
        for _ in range(repeats):
            t0 = time.perf_counter()
            grad = R.mri_AtA(xk, mask) - R.mri_At(y, mask)          # data_fidelity.py:335-336
            z = xk - STEPSIZE * grad                                  # pgd.py:137-139
            xk = R.drunet_forward(z, SIGMA_DEN, sd)                   # prior.py:99-109 -> drunet.py:212-263
            best = min(best, time.perf_counter() - t0)
    return best


def best_cpu_threads(sample: int):
    """the reference user would run torch's default (= all cores); MKL-DNN convolutions on a handful of images do not
    scale to 100+ threads, so the CPU arm reports the BEST of a few thread counts (the count used is reported)"""
    ncpu = os.cpu_count() or 1
    cands = sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 32), min(ncpu, 16)}, reverse=True)
    best = (float("inf"), ncpu)
    for th in cands:
        t = cpu_pgd_iteration_seconds(sample, 1, th)
        if t < best[0]:
            best = (t, th)
    return best[1]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 4
    threads = best_cpu_threads(sample)
    # warm-up + K timed "steps", each a bounded sample (one iteration on `sample` images)
    for _ in range(max(args.warmup, 0)):
        cpu_pgd_iteration_seconds(sample, 1, threads)
    ts = [cpu_pgd_iteration_seconds(sample, 1, threads) for _ in range(max(args.steps, 1))]
    t = sum(ts) / len(ts)
    per_img = t / sample
    value = 1.0 / (per_img * BATCH)  # every rank's shard is 64 images; the CPU arm runs them one shard after another
    value_job = value  # whole job at N GPUs = N shards on the same host: N x the work, N x the time
    line = {
        "impl": "reference", "metric": "pnp_pgd_iterations_per_s", "value": value_job, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value_job, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MRI 4x Cartesian-mask 256x256, PnP-PGD + DRUNet, batch=64 per GPU", "global_batch": BATCH * args.gpus,
                   "parallelism": f"dp{args.gpus}"},
        "cpu_baseline": {"value": value_job, "unit": "it/s", "cores": threads, "kind": "port",
                         "sample": f"one PnP-PGD iteration of the oracle (torch-CPU restatement of the reference path) on {sample} "
                                   f"of the 64 images per step, scaled linearly per image; best of thread counts, {threads} used"},
        "e2e": {"value": value_job, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def time_cuda(fn, iters: int, warmup: int = 3) -> float:
    """average milliseconds per call, CUDA events on the current stream"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def graph_time(calls, replays: int = 20) -> float:
    """ms per call: the calls (one per operand set) are captured into ONE CUDA graph and replayed"""
    for c in calls:
        c()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for c in calls:
            c()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        for c in calls:
            c()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * len(calls))


def run_b200(args):
    import torch.distributed as dist

    import deepinv_b200 as dinv
    from deepinv_b200.optim import L2, PGD, PnP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (deepinv_b200 has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()

    # ---- synthetic inputs (per-rank shard of the global batch) -----------------------------------
    torch.manual_seed(0)
    den = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None, precision=args.precision).to(dev).eval()
    gen = torch.Generator().manual_seed(1234 + rank)
    x_true = torch.randn(BATCH, 2, H, W, generator=gen)
    mask = cartesian_mask(BATCH, H, W, ACCEL, seed=rank)
    physics = dinv.physics.MRI(mask=mask.to(dev), img_size=(2, H, W), device=dev)
    x_pin = x_true.pin_memory()
    with torch.no_grad():
        y = physics.A(x_pin.to(dev, non_blocking=True))
    y_pin = y.cpu().pin_memory()
    algo = PGD(data_fidelity=L2(), prior=PnP(den), stepsize=STEPSIZE, sigma_denoiser=SIGMA_DEN, max_iter=args.steps,
               early_stop=False)

    lib = dinv.get_lib()

    def iteration(X, it):
        return algo.single_iteration(X, it, y, physics)

    graphed = None
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()  # from before the warm-up: the sampler needs ~100 ms to deliver its first line
    with torch.no_grad():
        X = algo.init_iterate_fn(y, physics)
        for it in range(args.warmup):
            X = iteration(X, it)
        if not args.no_graph:
            try:  # replay the same public-API iteration from a CUDA graph (no Python / ctypes launch path in the loop)
                from deepinv_b200.optim import GraphedIteration

                graphed = GraphedIteration(algo, y, physics, X=X)
                graphed.run(args.warmup)
            except Exception as exc:  # noqa: BLE001
                print(f"bench.py: CUDA-graph capture unavailable ({exc}); timing the eager loop", file=sys.stderr)
                graphed = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches0 = lib.dinvk_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_w0 = ClockSampler.now()
        e0.record()
        if graphed is not None:
            x_hat = graphed.run(args.steps)
        else:
            for it in range(args.steps):
                X = iteration(X, it)
            x_hat = X["est"][0]
        if world > 1:  # the only collective of the path: gather the final reconstructions
            gathered = torch.empty(world * BATCH, 2, H, W, device=dev)
            dist.all_gather_into_tensor(gathered, x_hat.contiguous())
        e1.record()
        torch.cuda.synchronize()
        t_w1 = ClockSampler.now()
        ms_total = e0.elapsed_time(e1)
        launches = lib.dinvk_launch_count() - launches0
        if graphed is not None:
            launches = graphed.launches_per_step * args.steps
        windows = [("timed region", t_w0, t_w1)]
        if rank == 0 and ms_total < 400.0:
            # the timed region is shorter than a handful of sampler periods: replay the SAME workload (untimed) for ~1 s
            # right away, back to back with the timed region, and read the clocks / throttle reasons from that window
            n_probe = max(1, int(1000.0 / max(ms_total / args.steps, 1e-3)))
            t_p0 = ClockSampler.now()
            if graphed is not None:
                graphed.run(n_probe)
            else:
                Xp = X
                for it in range(n_probe):
                    Xp = iteration(Xp, it)
            torch.cuda.synchronize()
            windows.append(("same workload replayed (untimed) for ~1 s right after the timed region", t_p0, ClockSampler.now()))
        clk = clocks.stop(windows) if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms_total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        ms_step = ms_total / args.steps
        value = world * 1000.0 / ms_step

        # ---- e2e: host buffers in, host buffer out, every step ------------------------------------
        xh = x_hat.cpu().pin_memory()
        out_pin = torch.empty(xh.shape, dtype=xh.dtype, pin_memory=True)  # (empty_like does not inherit pinned-ness)

        def e2e_step():
            xd = xh.to(dev, non_blocking=True)
            yd = y_pin.to(dev, non_blocking=True)
            Xn = algo.single_iteration({"est": (xd, xd), "aty": None}, 0, yd, physics)
            out_pin.copy_(Xn["est"][0], non_blocking=True)

        n_e2e = max(4, min(args.steps, 10))
        e2e_mode = "eager, one stream"
        pipe = None
        if not args.no_graph:
            try:  # uploads / graph replay / downloads on three streams, two device slots (public API: HostStreamedIteration)
                from deepinv_b200.optim import HostStreamedIteration

                pipe = HostStreamedIteration(algo, physics, xh, y_pin, dev)
                e2e_mode = "3-stream pipeline over 2 device slots, CUDA-graph compute"
            except Exception as exc:  # noqa: BLE001
                print(f"bench.py: host-streamed pipeline unavailable ({exc}); timing the eager e2e step", file=sys.stderr)
        if pipe is not None:
            def e2e_run(n):
                for _ in range(n):
                    pipe.submit(xh, y_pin, out_pin)
                pipe.drain()

            e2e_run(2)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            e2e_run(n_e2e)
            f1.record()
            torch.cuda.synchronize()
            ms_e2e = f0.elapsed_time(f1) / n_e2e
        else:
            ms_e2e = time_cuda(e2e_step, n_e2e, warmup=1)
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        h2d = xh.numel() * 4 + y_pin.numel() * 4
        d2h = out_pin.numel() * 4

        # ---- roofline of the dominant kernel + per-operator HBM fractions (rank 0) -------------------
        roof, ops_report, den_report = None, None, None
        if rank == 0:
            from deepinv_b200 import ops as dops

            z = x_hat
            ms_den = time_cuda(lambda: den(z, SIGMA_DEN), max(2, min(args.steps, 5)), warmup=1)
            tflops = DRUNET_GFLOP_PER_IMAGE * BATCH / ms_den  # GFLOP / ms = TFLOP/s
            den_report = {"what": "whole DRUNet forward (%s path, 64 conv launches)" % args.precision, "ms": ms_den,
                          "algorithmic_gflop": DRUNET_GFLOP_PER_IMAGE * BATCH, "TFLOPs": tflops,
                          "frac_of_sustained_bf16_peak": tflops / peaks["bf16_tflops_sustained"]}
            if args.precision == "bf16":
                # the dominant kernel of the step (largest share of the ncu launch list, profiles/): the 64->64 3x3 body
                # convolution at full resolution (ResBlock form: + residual), tcgen05 implicit GEMM.  One launch per call;
                # its 0.5 GB input and 0.5 GB output exceed the 126 MB L2, so every launch streams from HBM.
                C = 64
                xa = torch.randn(BATCH, H, W, C, device=dev).to(torch.bfloat16)
                ra = torch.randn(BATCH, H, W, C, device=dev).to(torch.bfloat16)
                wa = (torch.randn(C, 9 * C, device=dev) / (3 * C ** 0.5)).to(torch.bfloat16)
                ms_k = time_cuda(lambda: dops.conv3x3_bf16(xa, wa, res=ra), 20, warmup=3)
                gflop_k = 2.0 * BATCH * H * W * C * 9 * C / 1e9
                traffic = None
                tp = ROOT / "profiles" / "top_kernel_traffic.json"  # dram read+write bytes per launch from `ncu --set full`
                if tp.exists():
                    traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
                roof = {"bound": "tensor", "kernel": "conv_tc_halo_kernel<64,...>: 3x3 conv 64->64, 64x256x256, bf16 -> fp32 TMEM, +residual",
                        "achieved": gflop_k / ms_k, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": gflop_k / ms_k / peaks["bf16_tflops"], "traffic": traffic,
                        "peak_source": peaks["source"] + " bf16 burst (kernel timed alone, 20 back-to-back launches)",
                        "us_per_launch": ms_k * 1e3, "algorithmic_gflop_per_launch": gflop_k,
                        "algorithmic_bytes_per_launch": 3 * BATCH * H * W * C * 2 + C * 9 * C * 2}
                del xa, ra, wa
            else:
                roof = {"bound": "tensor", "kernel": "DRUNet convolutions (fp32 SIMT path, whole forward)", "achieved": tflops,
                        "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tflops / peaks["bf16_tflops_sustained"],
                        "traffic": None, "peak_source": peaks["source"] + " bf16 sustained"}
            # operators: each one captured into a CUDA graph that walks SETS disjoint operand sets (> L2 in total), so
            # every call streams its operands from HBM and no Python / ctypes launch cost is in the timed region
            SETS = 6
            img_mb = BATCH * 2 * H * W * 4 / 1e6
            gen2 = torch.Generator(device=dev).manual_seed(7)
            sets = []
            for i in range(SETS):
                xs = torch.randn(BATCH, 2, H, W, device=dev, generator=gen2)
                ms_ = cartesian_mask(BATCH, H, W, ACCEL, seed=100 + i).to(dev)
                ps = dinv.physics.MRI(mask=ms_, img_size=(2, H, W), device=dev)
                ys = ps.A(xs)
                sets.append((ps, xs, ys, ps.A_adjoint(ys)))
            cases = [
                ("MRI.A (2-D FFT + mask)", lambda p, x, y, aty: p.A(x), 2 * img_mb),
                ("MRI.A_adjoint (mask + 2-D iFFT)", lambda p, x, y, aty: p.A_adjoint(y), 2 * img_mb),
                ("PGD data step x-g(AtAx-Aty), line mask (1 pass)", lambda p, x, y, aty: p.normal_step(x, aty, STEPSIZE), 3 * img_mb),
                ("MRI.prox_l2, line mask (1 pass)", lambda p, x, y, aty: p.prox_l2(x, y, 1.0), 3 * img_mb),
            ]
            ops_report = []
            for name, fn, mb in cases:
                ms = graph_time([(lambda st=st, fn=fn: fn(*st)) for st in sets])
                gbs = mb / ms  # MB/ms = GB/s
                ops_report.append({"op": name, "ms": ms, "algorithmic_MB": mb, "GBps": gbs, "frac_hbm": gbs / peaks["hbm_gbs"],
                                   "timing": f"CUDA graph over {SETS} disjoint operand sets (> L2), 20 replays"})
            del sets

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = best_cpu_threads(4)
            t = cpu_pgd_iteration_seconds(4, 2, threads)
            cpu = {"value": 1.0 / (t / 4 * BATCH), "unit": "it/s", "cores": threads, "kind": "port",
                   "sample": "one PnP-PGD iteration of the oracle on 4 of the 64 images (best of 2, best of thread counts), "
                             "scaled linearly per image"}
        line = {
            "metric": "pnp_pgd_iterations_per_s", "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 denoiser GEMMs (fp32 accumulate) + f32 operators" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": "MRI 4x Cartesian-mask 256x256, PnP-PGD + DRUNet, batch=64 per GPU",
                       "global_batch": BATCH * world, "parallelism": f"dp{world}", "denoiser_precision": args.precision, "cuda_graph": graphed is not None,
                       "l2_policy": "per-step working set (>= 1 GB of activations) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": world * 1000.0 / ms_e2e, "unit": "it/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e, "mode": e2e_mode},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "denoiser": den_report,
            "operators": ops_report,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DINVK_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager Python loop instead of CUDA-graph replays")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

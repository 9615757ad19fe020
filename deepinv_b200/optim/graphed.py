"""CUDA-graph replay of one optimiser iteration.

The reference's Python loop (fixed_point.py:324-359) costs tens of microseconds of host work per iteration and ~70
kernel launches through ctypes here; both disappear when the iteration is captured once and replayed.  The capture
goes through the public `BaseOptim.single_iteration`, so the replayed work is exactly the eager iteration (same
kernels, same order).  Only iteration-independent algorithms can be captured (PGD, HQS, ADMM with constant parameters;
FISTA's momentum depends on the iteration counter).
"""
from __future__ import annotations

import torch


class GraphedIteration:
    def __init__(self, algo, y: torch.Tensor, physics, X: dict | None = None, it: int = 0, warmup: int = 2):
        if any(len(v) > 1 for v in algo.init_params_algo.values()):
            raise ValueError("per-iteration parameter schedules cannot be captured in a single graph")
        self.algo, self.y, self.physics = algo, y, physics
        with torch.no_grad():
            X = algo.init_iterate_fn(y, physics) if X is None else X
            self.x = X["est"][0].clone()
            self.z = X["est"][1].clone()
            self.aty = X.get("aty")
            side = torch.cuda.Stream(device=y.device)
            side.wait_stream(torch.cuda.current_stream(y.device))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._body(it)
            torch.cuda.current_stream(y.device).wait_stream(side)
            torch.cuda.synchronize(y.device)
            from .._lib import launch_count

            n0 = launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                xn, zn = self._body(it)
                self.x.copy_(xn)
                self.z.copy_(zn)
            self.launches_per_step = launch_count() - n0  # libdinvk kernels recorded in the graph

    def _body(self, it):
        Xn = self.algo.single_iteration({"est": (self.x, self.z), "aty": self.aty}, it, self.y, self.physics)
        return Xn["est"][0], Xn["est"][1]

    def load(self, x: torch.Tensor, z: torch.Tensor | None = None) -> None:
        self.x.copy_(x)
        self.z.copy_(x if z is None else z)

    def step(self) -> torch.Tensor:
        """one iteration in place; returns the (static) iterate tensor"""
        self.graph.replay()
        return self.x

    def run(self, n: int) -> torch.Tensor:
        for _ in range(n):
            self.graph.replay()
        return self.x


class GraphedSolve:
    """A WHOLE reconstruction `algo(y, physics)` — every iteration, per-iteration schedules (DPIR's sigma / stepsize) baked
    in as constants — captured once as one CUDA graph and replayed per measurement: `solve(y)` costs one copy of y into the
    static input buffer plus one graph launch instead of max_iter x ~70 ctypes launches.  Requires a run without host
    synchronisation: `early_stop=False`, no metrics, closed-form data steps (MRI / BlurFFT prox, fused gradient step); the
    CG-based prox polls a device flag from the host and cannot be captured."""

    def __init__(self, algo, y: torch.Tensor, physics, warmup: int = 1):
        if getattr(algo, "early_stop", False):
            raise ValueError("early_stop=True needs a host decision per iteration and cannot be captured")
        self.algo, self.physics = algo, physics
        self.y = y.clone()
        dev = y.device
        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    out = algo(self.y, physics)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            from .._lib import launch_count

            n0 = launch_count()
            self.out = torch.empty_like(out)
            self.y.add_(0)  # bump the version: operators that memoise A^T y per (buffer, version) recompute it INSIDE the graph
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out.copy_(algo(self.y, physics))
            self.launches = launch_count() - n0

    def solve(self, y: torch.Tensor) -> torch.Tensor:
        """reconstruction of `y` (same shape as the capture's); the returned tensor is the static output buffer"""
        self.y.copy_(y)
        self.graph.replay()
        return self.out


class HostStreamedIteration:
    """A stream of independent single-iteration requests whose inputs AND outputs live in pinned host memory.

    Request k uploads its iterate x_k and measurement y_k (host -> device), runs one iteration of `algo`
    (`A^T y_k` included — nothing is cached across requests) and downloads the new iterate.  Uploads, compute and
    downloads run on three CUDA streams over two device slots, so the PCIe traffic of request k+1 / k-1 overlaps the
    compute of request k; every request still moves all of its bytes over PCIe.  Compute is a CUDA-graph replay of the
    public `single_iteration` (one graph per slot, bound to that slot's buffers).
    """

    def __init__(self, algo, physics, x_host: torch.Tensor, y_host: torch.Tensor, device, slots: int = 2, it: int = 0):
        if any(len(v) > 1 for v in algo.init_params_algo.values()):
            raise ValueError("per-iteration parameter schedules cannot be captured in a single graph")
        self.device = device
        self.n = 0
        self.up, self.down = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        self.slots = []
        with torch.no_grad():
            for _ in range(slots):
                s = {
                    "x": torch.empty(x_host.shape, dtype=x_host.dtype, device=device),
                    "y": torch.empty(y_host.shape, dtype=y_host.dtype, device=device),
                    "uploaded": torch.cuda.Event(), "computed": torch.cuda.Event(), "downloaded": torch.cuda.Event(),
                }
                s["x"].copy_(x_host)
                s["y"].copy_(y_host)
                body = lambda s=s: algo.single_iteration({"est": (s["x"], s["x"]), "aty": None}, it, s["y"], physics)["est"][0]
                side = torch.cuda.Stream(device=device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    s["out"] = torch.empty_like(body())
                torch.cuda.current_stream(device).wait_stream(side)
                torch.cuda.synchronize(device)
                s["graph"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(s["graph"]):
                    s["out"].copy_(body())
                self.slots.append(s)

    def submit(self, x_host: torch.Tensor, y_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """enqueue one request (asynchronous; `out_host` is valid after `drain()`)"""
        s = self.slots[self.n % len(self.slots)]
        first_use = self.n < len(self.slots)
        self.n += 1
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.up):
            if not first_use:
                self.up.wait_event(s["computed"])  # the slot's previous request no longer reads x / y
            s["x"].copy_(x_host, non_blocking=True)
            s["y"].copy_(y_host, non_blocking=True)
            s["uploaded"].record(self.up)
        cur.wait_event(s["uploaded"])
        if not first_use:
            cur.wait_event(s["downloaded"])  # the slot's previous result has left the device
        s["graph"].replay()
        s["computed"].record(cur)
        with torch.cuda.stream(self.down):
            self.down.wait_event(s["computed"])
            out_host.copy_(s["out"], non_blocking=True)
            s["downloaded"].record(self.down)

    def drain(self) -> None:
        """make the current stream wait for every enqueued upload / download"""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.up)
        cur.wait_stream(self.down)

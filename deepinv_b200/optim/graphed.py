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

"""CUDA-graph replay and the host-streamed request pipeline must reproduce the eager public-API iteration exactly
(same kernels in the same order: bitwise equality is the bar)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _setup(dev, precision="fp32", B=4, H=64, W=64):
    import deepinv_b200 as dinv
    from deepinv_b200.optim import L2, PGD, PnP

    torch.manual_seed(0)
    den = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None, precision=precision).to(dev).eval()
    cols = (torch.rand(B, 1, 1, W) > 0.7).float().expand(B, 2, H, W).contiguous()
    physics = dinv.physics.MRI(mask=cols.to(dev), img_size=(2, H, W), device=dev)
    algo = PGD(data_fidelity=L2(), prior=PnP(den), stepsize=1.0, sigma_denoiser=0.05, max_iter=4, early_stop=False)
    return physics, algo


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_iteration_equals_eager(precision, dev):
    from deepinv_b200.optim import GraphedIteration

    physics, algo = _setup(dev, precision)
    with torch.no_grad():
        y = physics.A(torch.randn(4, 2, 64, 64, device=dev))
        X = algo.init_iterate_fn(y, physics)
        g = GraphedIteration(algo, y, physics, X={k: (tuple(t.clone() for t in v) if isinstance(v, tuple) else v) for k, v in X.items()})
        assert g.launches_per_step > 0
        for it in range(3):
            X = algo.single_iteration(X, it, y, physics)
        xg = g.run(3)
        torch.cuda.synchronize()
        assert torch.equal(xg, X["est"][0])


def test_host_streamed_requests(dev):
    from deepinv_b200.optim import HostStreamedIteration

    physics, algo = _setup(dev, "bf16")
    gen = torch.Generator().manual_seed(3)
    n = 5  # more requests than slots: exercises slot reuse and the upload/compute/download ordering
    xs = [torch.randn(4, 2, 64, 64, generator=gen).pin_memory() for _ in range(n)]
    with torch.no_grad():
        ys = [physics.A(torch.randn(4, 2, 64, 64, generator=gen).to(dev)).cpu().pin_memory() for _ in range(n)]
        outs = [torch.empty(4, 2, 64, 64).pin_memory() for _ in range(n)]
        pipe = HostStreamedIteration(algo, physics, xs[0], ys[0], dev)
        for k in range(n):
            pipe.submit(xs[k], ys[k], outs[k])
        pipe.drain()
        torch.cuda.synchronize()
        for k in range(n):
            xd, yd = xs[k].to(dev), ys[k].to(dev)
            ref = algo.single_iteration({"est": (xd, xd), "aty": None}, 0, yd, physics)["est"][0]
            assert torch.equal(outs[k].to(dev), ref), f"request {k}"

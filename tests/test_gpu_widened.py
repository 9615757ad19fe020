"""GPU parity of the rows widened from SURVEY §8(f) / §8(b) (DRS, GD, DPIR, DEQ, training closure, DynamicMRI / SequentialMRI,
3-D MRI, Downsampling, compose / stack, device-side generators, whole-run CUDA graphs) through the package's public API on
cuda:0, against golden vectors of the real reference.  Same cases as the CPU host-logic tests (tests/parity_cases.py); kept in
a file that sorts after the established GPU suites."""
import pytest
import torch

import parity_cases as P
from conftest import golden_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", golden_names("dynmri_") + golden_names("seqmri_"))
def test_dynamic_mri(name, dev):
    P.case_dynamic_mri(name, dev)


@pytest.mark.parametrize("name", golden_names("fan_"))
def test_fanbeam(name, dev):
    P.case_fanbeam(name, dev)


@pytest.mark.parametrize("name", golden_names("down_"))
def test_downsampling(name, dev):
    P.case_downsampling(name, dev)


def test_combine(dev):
    P.case_combine(dev)


def test_mri_3d(dev):
    P.case_mri_3d(dev)


def test_anderson(dev):
    P.case_anderson(dev)


def test_diffpir(dev):
    P.case_diffpir(dev)


def test_filters(dev):
    P.case_filters(dev)


def test_drs_gd_dpir(dev):
    P.case_drs_gd_dpir(dev)


def test_train_deq_explicit(dev):
    P.case_train_deq_explicit(dev)


def test_train_unfolded(dev):
    P.case_train_unfolded(dev)


def test_training_loop_reduces_loss(dev):
    """end to end: a few Adam steps on an unfolded PGD model (trainable stepsizes + DRUNet weights) whose every forward
    and backward op is a libdinvk launch; the supervised loss must go down"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import L2, PnP
    from deepinv_b200.unfolded import unfolded_builder

    torch.manual_seed(0)
    B, H, W = 4, 32, 32
    x = torch.randn(B, 2, H, W, device=dev) * 0.5
    cols = (torch.rand(B, 1, 1, W) > 0.6).float().expand(B, 2, H, W).contiguous().to(dev)
    phys = dinv.physics.MRI(mask=cols, img_size=(2, H, W), device=dev)
    with torch.no_grad():
        y = phys.A(x)
    den = dinv.models.DRUNet(in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=1, pretrained=None, device=dev).train()
    model = unfolded_builder("PGD", params_algo={"stepsize": [1.0, 1.0, 1.0], "g_param": 0.05, "lambda": 1.0},
                             trainable_params=["stepsize"], data_fidelity=L2(), prior=PnP(den), max_iter=3).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    n0 = dinv.launch_count()
    losses = []
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        loss = ((model(y, phys) - x) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert dinv.launch_count() - n0 > 8 * 3 * 20
    # (the same loop in plain torch on the oracle: 0.2526 -> 0.226 in 8 steps)
    assert all(b < a for a, b in zip(losses, losses[1:])) and losses[-1] < 0.95 * losses[0], losses



def test_mask_generators_cuda_feed_mri(dev):
    from test_generators import _check

    _check("cuda:0")
    import deepinv_b200 as dinv
    from deepinv_b200.physics.generator import RandomMaskGenerator

    dev = torch.device("cuda:0")
    gen = RandomMaskGenerator((2, 64, 64), acceleration=4, device=dev, rng=torch.Generator(device=dev).manual_seed(0))
    phys = dinv.physics.MRI(img_size=(2, 64, 64), device=dev)
    x = torch.randn(4, 2, 64, 64, device=dev)
    y = phys(x, **gen.step(4))  # mask generated on the device, stored by the forward call (forward.py:249-276)
    assert phys.mask.shape == (4, 2, 64, 64) and float((y != 0).float().mean()) == pytest.approx(0.25, abs=0.01)


def test_graphed_solve_dpir(dev):
    """a whole DPIR reconstruction (8 HQS iterations with their sigma / stepsize schedule) as ONE graph: replays on new
    measurements reproduce the eager runs bit for bit"""
    import deepinv_b200 as dinv
    from deepinv_b200.optim import DPIR, GraphedSolve

    torch.manual_seed(0)
    B, H, W = 2, 64, 64
    den = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None, precision="bf16").to(dev).eval()
    cols = (torch.rand(B, 1, 1, W) > 0.7).float().expand(B, 2, H, W).contiguous()
    physics = dinv.physics.MRI(mask=cols.to(dev), img_size=(2, H, W), device=dev)
    algo = DPIR(sigma=0.05, denoiser=den, device=dev)
    with torch.no_grad():
        ys = [physics.A(torch.randn(B, 2, H, W, device=dev)) for _ in range(3)]
        g = GraphedSolve(algo, ys[0], physics)
        assert g.launches > 8 * 60
        for y in ys[::-1]:
            want = algo(y.clone(), physics)
            got = g.solve(y)
            torch.cuda.synchronize()
            assert torch.equal(got, want)

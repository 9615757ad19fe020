"""GPU parity against the reference's golden vectors through the package's public API (C-ABI kernels on
cuda:0).  Same cases as tests/test_host_logic_emul.py, all fixtures.  Tolerance 1e-5 relative L2 (fp32),
exact zero pattern for masked k-space; CG-based results 1e-4 (see parity_cases.py)."""
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


@pytest.mark.parametrize("name", golden_names("mri_"))
def test_mri(name, dev):
    P.case_mri(name, dev)


@pytest.mark.parametrize("name", golden_names("mcmri_"))
def test_multicoil(name, dev):
    P.case_multicoil(name, dev)


@pytest.mark.parametrize("name", [n for n in golden_names("tomo_") if "norm" not in n])
def test_tomography(name, dev):
    P.case_tomography(name, dev)


def test_tomography_normalised(dev):
    P.case_tomography_normalised(dev)


@pytest.mark.parametrize("name", [n for n in golden_names("blur_") if "prox" not in n])
def test_blur(name, dev):
    P.case_blur(name, dev)


def test_blur_cg(dev):
    P.case_blur_cg(dev)


@pytest.mark.parametrize("name", golden_names("blurfft_"))
def test_blurfft(name, dev):
    P.case_blurfft(name, dev)


def test_drunet(dev):
    P.case_drunet(dev)


def test_dncnn(dev):
    P.case_dncnn(dev)


def test_pnp_mri(dev):
    P.case_pnp_mri(dev)


def test_pnp_blur_admm(dev):
    P.case_pnp_blur_admm(dev)


def test_ddrm(dev):
    P.case_ddrm(dev)


def test_library_loaded_is_in_tree(dev):
    import deepinv_b200 as dinv

    assert dinv._lib.lib_path().name == "libdinvk.so" and dinv._lib.lib_path().exists()
    n0 = dinv.launch_count()
    dinv.physics.MRI(img_size=(2, 8, 8), device=dev).A(torch.randn(1, 2, 8, 8, device=dev))
    assert dinv.launch_count() > n0

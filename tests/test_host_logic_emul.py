"""Host-side logic without a GPU: the package's own Python code (physics classes, optimisers, denoisers,
sampler) driving the EMULATED kernels (tests/emul) on CPU tensors, checked against the reference's golden
vectors.  The emulated backend is injected by monkeypatching three private hooks of deepinv_b200.ops inside
this test session only; the package itself has no CPU switch (see tests/test_abi.py::test_no_cpu_fallback)."""
import pytest
import torch

import parity_cases as P
from conftest import golden_names


@pytest.fixture(autouse=True)
def emul_backend(monkeypatch):
    from emul_util import emul_lib

    from deepinv_b200 import ops

    lib = emul_lib()

    def check(rc):
        assert rc == 0, lib.dinvk_last_error()

    monkeypatch.setattr(ops, "_require_cuda", lambda *ts: torch.device("cpu"))
    monkeypatch.setattr(ops, "_stream", lambda dev: None)
    monkeypatch.setattr(ops, "get_lib", lambda: lib)
    monkeypatch.setattr(ops, "check", check)
    ops._ws_cache.clear()
    yield
    ops._ws_cache.clear()


DEV = torch.device("cpu")


@pytest.mark.parametrize("name", golden_names("mri_"))
def test_mri(name):
    P.case_mri(name, DEV)


@pytest.mark.parametrize("name", golden_names("mcmri_"))
def test_multicoil(name):
    P.case_multicoil(name, DEV)


@pytest.mark.parametrize("name", ["tomo_16_a8", "tomo_24_a10_circle"])
def test_tomography(name):
    P.case_tomography(name, DEV)


def test_tomography_normalised():
    P.case_tomography_normalised(DEV)


@pytest.mark.parametrize("name", ["blur_3x3_valid", "blur_4x4_circular", "blur_5x3_replicate", "blur_6x5_reflect",
                                  "blur_4x4_constant", "blur_5x5_perbc_reflect"])
def test_blur(name):
    P.case_blur(name, DEV)


def test_blur_cg():
    P.case_blur_cg(DEV)


@pytest.mark.parametrize("name", ["blurfft_18x20", "blurfft_15x16_odd"])
def test_blurfft(name):
    P.case_blurfft(name, DEV)


def test_drunet():
    P.case_drunet(DEV)


def test_dncnn():
    P.case_dncnn(DEV)


def test_pnp_mri():
    P.case_pnp_mri(DEV)


def test_pnp_blur_admm():
    P.case_pnp_blur_admm(DEV)


def test_ddrm():
    P.case_ddrm(DEV)

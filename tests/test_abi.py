"""C-ABI surface checks that need no GPU: the shared library builds, loads, and exports every symbol
include/dinvk.h declares; the ctypes table matches the header; the product refuses CPU tensors."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    txt = (ROOT / "include" / "dinvk.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dinvk_[a-z0-9_]+)\s*\(", txt)))


def test_header_matches_ctypes_table():
    from deepinv_b200 import _ffi

    assert header_symbols() == _ffi.declared_symbols()


def test_library_builds_loads_and_exports_everything():
    from deepinv_b200 import _ffi
    from deepinv_b200.build import build

    lib = ctypes.CDLL(str(build()))
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"libdinvk.so does not export: {missing}"
    _ffi.bind(lib, required=True)
    assert lib.dinvk_version() >= 100
    assert lib.dinvk_spectral_workspace_bytes(2, 8, 8) >= 2 * 2 * 64 * 8


def test_argument_errors_are_codes_not_exceptions():
    from deepinv_b200 import _ffi
    from deepinv_b200.build import build

    lib = _ffi.bind(ctypes.CDLL(str(build())), required=False)
    assert lib.dinvk_spectral(None, None, 0, None) == 1  # DINVK_EINVAL
    assert b"null" in lib.dinvk_last_error()
    assert lib.dinvk_axpbypcz(None, None, 1.0, None, 0.0, None, 0.0, 4, None) == 1


def test_no_cpu_fallback():
    import deepinv_b200 as dinv

    phys = dinv.physics.MRI(img_size=(2, 8, 8))
    with pytest.raises(dinv.DinvkError):
        phys.A(torch.randn(1, 2, 8, 8))
    with pytest.raises(dinv.DinvkError), torch.no_grad():
        dinv.models.DnCNN(in_channels=1, out_channels=1, pretrained=None)(torch.randn(1, 1, 8, 8))


def test_state_dict_compatibility_with_reference_layout():
    """the reference's DRUNet/DnCNN key names and shapes (tests/golden fixtures hold real state_dicts)"""
    import deepinv_b200 as dinv
    from conftest import load_golden

    sd = load_golden("drunet_tiny")["sd"]
    m = dinv.models.DRUNet(in_channels=2, out_channels=2, nc=(8, 16, 32, 64), nb=2, pretrained=None)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    sd = load_golden("dncnn_tiny")["sd"]
    m = dinv.models.DnCNN(in_channels=1, out_channels=1, depth=5, nf=8, pretrained=None)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    full = dinv.models.DRUNet(in_channels=2, out_channels=2, pretrained=None)
    assert sum(p.numel() for p in full.parameters()) == 32_639_808


def test_public_api_surface():
    """the names README.md lists under the reference's module layout exist"""
    import deepinv_b200 as d

    table = {
        d.physics: "Physics LinearPhysics DecomposablePhysics MRI MultiCoilMRI DynamicMRI SequentialMRI Tomography Blur BlurFFT "
                   "Downsampling Denoising Inpainting compose stack TensorList GaussianNoise",
        d.physics.functional: "gaussian_blur bilinear_filter bicubic_filter sinc_filter kaiser_window",
        d.physics.generator: "RandomMaskGenerator GaussianMaskGenerator EquispacedMaskGenerator MotionBlurGenerator",
        d.optim: "L2 PnP RED Tikhonov ZeroPrior PGD FISTA ADMM HQS DRS GD PDCP DPIR BaseOptim optim_builder create_iterator least_squares "
                 "conjugate_gradient bicgstab lsqr minres GraphedIteration GraphedSolve HostStreamedIteration DEQConfig "
                 "AndersonAccelerationConfig BacktrackingConfig",
        d.unfolded: "unfolded_builder BaseUnfold DEQ_builder BaseDEQ",
        d.models: "DRUNet DnCNN",
        d.sampling: "DDRM DiffPIR",
    }
    missing = [f"{m.__name__}.{n}" for m, names in table.items() for n in names.split() if not hasattr(m, n)]
    assert not missing, missing

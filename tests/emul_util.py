"""helpers for the TEST-ONLY host emulation of the SIMT kernels (tests/emul)"""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "emul"))

_lib = None


def emul_lib():
    global _lib
    if _lib is None:
        from build_emul import build

        from deepinv_b200 import _ffi

        _lib = _ffi.bind(C.CDLL(str(build())), required=False)
    return _lib


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def spectral_emul(p0, H, W, fwd, inv, gmode=0, mask=None, mask_strides=(0, 0, 0), centered=True, a0=1.0, p1=None, a1=0.0,
                  c=0.0, c_batch=None, q0=None, e1=0.0, q1=None, e2=0.0, e0=1.0, ncoil=0, coil_mode=0, coil_maps=None,
                  out_shape=None):
    from deepinv_b200 import _ffi

    lib = emul_lib()
    p0 = p0.contiguous()
    p1, q0, q1 = (None if t is None else t.contiguous() for t in (p1, q0, q1))
    nimg = p0.shape[0] * (ncoil if ncoil > 1 else 1) if ncoil > 1 else p0.numel() // (2 * H * W)
    out = torch.zeros(out_shape if out_shape is not None else tuple(p0.shape), dtype=torch.float32)
    a = _ffi.SpectralArgs()
    a.B, a.H, a.W, a.fwd, a.inv, a.centered, a.gmode = nimg, H, W, int(fwd), int(inv), int(centered), gmode
    a.p0, a.p1, a.a0, a.a1 = p0.data_ptr(), (p1.data_ptr() if p1 is not None else None), a0, a1
    if mask is not None:
        a.mask = mask.data_ptr()
        a.mask_sb, a.mask_sc, a.mask_sh = mask_strides
    a.c = c
    a.c_batch = c_batch.data_ptr() if c_batch is not None else None
    a.q0, a.q1 = (q0.data_ptr() if q0 is not None else None), (q1.data_ptr() if q1 is not None else None)
    a.e0, a.e1, a.e2 = e0, e1, e2
    a.out = out.data_ptr()
    a.ncoil, a.coil_mode = (ncoil if ncoil > 1 else 0), coil_mode
    if coil_maps is not None:
        a.coil_maps = coil_maps.data_ptr()
        a.coil_sb = ncoil * H * W if coil_maps.shape[0] > 1 else 0
    nb = lib.dinvk_spectral_workspace_bytes(nimg, H, W)
    ws = torch.zeros(nb, dtype=torch.uint8)
    rc = lib.dinvk_spectral(C.byref(a), ptr(ws), nb, None)
    assert rc == 0, lib.dinvk_last_error()
    return out

"""TEST-ONLY: compile libdinvk's SIMT kernel sources for the host with g++ against cuda_emul.h.

The result (tests/emul/_build/libdinvk_emul.so) is loaded only by tests/test_emul_*.py through
ctypes with host (numpy) buffers.  The deepinv_b200 package never loads it.
"""
from __future__ import annotations

import hashlib
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "deepinv_b200" / "csrc"
OUT = HERE / "_build" / "libdinvk_emul.so"
# SIMT-only translation units (the tcgen05/TMA kernels cannot be emulated)
SOURCES = ["core.cu", "spectral.cu", "elementwise.cu", "radon.cu", "blur.cu", "conv_simt.cu"]


def build() -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    h = hashlib.sha256()
    for f in sorted(srcs + list(CSRC.glob("*.cuh")) + [HERE / "cuda_emul.h", ROOT / "include" / "dinvk.h", Path(__file__)]):
        h.update(f.read_bytes())
    stamp = OUT.with_suffix(".stamp")
    if OUT.exists() and stamp.exists() and stamp.read_text() == h.hexdigest():
        return OUT
    OUT.parent.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = OUT.parent / (s.stem + ".o")
        objs.append(o)
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-DDINVK_EMUL", "-x", "c++", "-I", str(HERE), "-I", str(ROOT / "include"),
               "-Wno-unknown-pragmas", "-c", str(s), "-o", str(o)]
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("emul build failed: " + " ".join(cmd))
    subprocess.run(["g++", "-shared", "-o", str(OUT), *map(str, objs), "-lpthread"], check=True)
    stamp.write_text(h.hexdigest())
    return OUT


if __name__ == "__main__":
    print(build())

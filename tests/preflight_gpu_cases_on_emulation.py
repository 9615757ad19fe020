#!/usr/bin/env python
"""Pre-flight for a GPU session: run EVERY parity case of the GPU suites (tests/parity_cases.py, all fixtures, full variants) on the
host emulation of the kernels (tests/emul) — about 3.5 minutes on 8 vCPUs, no GPU minutes.  The emulated backend is injected exactly
like in tests/test_host_logic_emul.py; this is test tooling, not a product path."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tests'))
sys.path.insert(0, str(ROOT))
import torch
from deepinv_b200 import ops
from emul_util import emul_lib
lib = emul_lib()
def check(rc): assert rc==0, lib.dinvk_last_error()
ops._require_cuda=lambda *ts: torch.device("cpu"); ops._stream=lambda dev: None; ops.get_lib=lambda: lib; ops.check=check
import parity_cases as P
from conftest import golden_names
D=torch.device("cpu"); t0=time.time()
for n in golden_names("mri_"): P.case_mri(n, D)
for n in golden_names("mcmri_"): P.case_multicoil(n, D)
for n in golden_names("dynmri_")+golden_names("seqmri_"): P.case_dynamic_mri(n, D)
for n in [n for n in golden_names("tomo_") if "norm" not in n]: P.case_tomography(n, D)
P.case_tomography_normalised(D)
for n in golden_names("fan_"): P.case_fanbeam(n, D)
for n in [n for n in golden_names("blur_") if "prox" not in n]: P.case_blur(n, D)
P.case_blur_cg(D)
for n in golden_names("blurfft_"): P.case_blurfft(n, D)
for n in golden_names("down_"): P.case_downsampling(n, D)
P.case_combine(D); P.case_mri_3d(D); P.case_filters(D); P.case_drunet(D); P.case_dncnn(D)
P.case_pnp_blur_admm(D); P.case_anderson(D); P.case_diffpir(D); P.case_train_deq_explicit(D)
print("all but training OK", round(time.time()-t0))
P.case_train_unfolded(D); print("training OK", round(time.time()-t0))

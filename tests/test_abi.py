"""The C-ABI library loads and exports every symbol include/aas_lmfb.h declares; host-only
entry points behave.  No GPU, no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from aas_enhancement_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "aas_lmfb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aas_(?:lmfb|l1)_[a-z_0-9]+)\s*\(", text)))


def test_exports_match_header(lib):
    from aas_enhancement_b200 import _lib
    names = _declared_functions()
    assert set(names) == set(_lib.EXPORTS)
    for name in names:
        assert getattr(lib, name) is not None


def test_abi_version_and_strerror(lib):
    assert lib.aas_lmfb_abi_version() == 1
    assert lib.aas_lmfb_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert b"aas_lmfb" in lib.aas_lmfb_strerror(code)


def test_plan_create_accepts_triangular_and_rejects_dense(lib):
    from aas_enhancement_b200 import MelPlan, slaney_mel_basis
    for n_mels in (40, 64, 80):
        assert MelPlan(slaney_mel_basis(n_mels=n_mels)).handle
    with pytest.raises(RuntimeError, match="not banded"):
        MelPlan(np.ones((40, 161), dtype=np.float32))
    with pytest.raises(RuntimeError, match="unsupported shape"):
        MelPlan(np.ones((40, 257), dtype=np.float32))


def test_workspace_bytes(lib):
    assert lib.aas_lmfb_workspace_bytes(30, 40, 601, 5) == 30 * 40 * 601 * 4
    assert lib.aas_lmfb_workspace_bytes(0, 40, 601, 5) == 0


def test_argument_errors_do_not_touch_the_gpu(lib):
    from aas_enhancement_b200 import MelPlan, slaney_mel_basis
    plan = MelPlan(slaney_mel_basis())
    rc = lib.aas_lmfb_forward(plan.handle, None, None, 1, 0, None, None, 0, 0, None, None, None,
                              10, 5, 0.0, None, None)
    assert rc == -1
    rc = lib.aas_lmfb_forward(plan.handle, 16, 16, 1, 0, 16, 16, 0, 0, 16, 16, 16, 10, 0xFF, 0.0, None, None)
    assert rc == -4
    rc = lib.aas_lmfb_forward(plan.handle, 18, 16, 1, 0, 16, 16, 0, 0, 16, 16, 16, 10, 5, 0.0, None, None)
    assert rc == -2


def test_python_front_end_refuses_cpu_tensors(lib):
    import torch
    from aas_enhancement_b200 import LMFBFrontEnd
    fe = LMFBFrontEnd()
    with pytest.raises(RuntimeError, match="CUDA-only"):
        fe(torch.zeros(1, 1600), torch.tensor([1600], dtype=torch.int32),
           torch.zeros(1, 161, 11), torch.zeros(1, 161, 11))


def test_default_mel_and_window_match_oracle(lib):
    from oracle import lmfb_oracle as orc
    from aas_enhancement_b200 import slaney_mel_basis, hamming_window
    assert np.abs(slaney_mel_basis() - orc.mel_filterbank()).max() < 1e-15
    assert np.abs(hamming_window() - orc.hamming_window()).max() == 0

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
    assert lib.aas_lmfb_abi_version() == 2
    assert lib.aas_lmfb_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert b"aas_lmfb" in lib.aas_lmfb_strerror(code)


def _info(lib, plan):
    m, c, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.aas_lmfb_plan_info(plan.handle, ctypes.byref(m), ctypes.byref(c), ctypes.byref(b)) == 0
    return m.value, c.value, b.value


def test_plan_create_accepts_any_basis(lib):
    """The reference applies any (M, F) matrix (model.py:167, :196): triangular filterbanks take the fast
    path, dense / re-ordered ones the generic path; only the shape is checked."""
    from aas_enhancement_b200 import MelPlan, slaney_mel_basis
    for n_mels in (40, 64, 80):
        assert _info(lib, MelPlan(slaney_mel_basis(n_mels=n_mels))) == (n_mels, 1, 1)
    assert _info(lib, MelPlan(np.ones((40, 161), dtype=np.float32))) == (40, 0, 0)
    assert _info(lib, MelPlan(slaney_mel_basis()[::-1].copy())) == (40, 0, 1)      # reversed order: not walkable, but still two adjacent rows per bin
    perm = slaney_mel_basis()[np.random.RandomState(0).permutation(40)]
    assert _info(lib, MelPlan(perm)) == (40, 0, 0)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        MelPlan(np.ones((40, 257), dtype=np.float32))


def test_workspace_bytes_and_tuning(lib):
    from aas_enhancement_b200 import MelPlan, slaney_mel_basis
    plan = MelPlan(slaney_mel_basis())
    assert lib.aas_lmfb_workspace_bytes(plan.handle, 30, 601, 5) == 30 * 40 * 601 * 4
    assert lib.aas_lmfb_workspace_bytes(plan.handle, 0, 601, 5) == 0
    dense = MelPlan(np.ones((40, 161), dtype=np.float32))
    assert lib.aas_lmfb_workspace_bytes(dense.handle, 30, 601, 5) == 30 * (40 + 162) * 601 * 4
    assert lib.aas_lmfb_plan_set_tuning(plan.handle, 4, 6, 1) == 0
    assert lib.aas_lmfb_plan_set_tuning(plan.handle, 7, 0, 0) == -4
    assert lib.aas_lmfb_plan_set_tuning(plan.handle, 0, 0, 0) == 0


def test_io_struct_matches_the_header(lib):
    """ctypes mirror of aas_lmfb_io: same field names, in the header's order."""
    from aas_enhancement_b200 import _lib
    text = open(os.path.join(ROOT, "include", "aas_lmfb.h")).read()
    body = text[text.index("typedef struct aas_lmfb_io {"):text.index("} aas_lmfb_io;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"([a-z_0-9]+)\s*;", body)
    assert names == [f[0] for f in _lib.IO._fields_]
    assert ctypes.sizeof(_lib.IO) == 8 * 4 + 8 * 24


def test_argument_errors_do_not_touch_the_gpu(lib):
    from aas_enhancement_b200 import MelPlan, slaney_mel_basis
    plan = MelPlan(slaney_mel_basis())
    rc = lib.aas_lmfb_forward(plan.handle, None, None, 1, 0, None, None, 0, 0, None, None, None,
                              10, 5, 0.0, None, None)
    assert rc == -1
    rc = lib.aas_lmfb_forward(plan.handle, 16, 16, 1, 0, 16, 16, 0, 0, 16, 16, 16, 10, 0xFF, 0.0, None, None)
    assert rc == -4
    rc = lib.aas_lmfb_forward(plan.handle, 18, 16, 1, 0, 16, 16, 0, 10, 16, 16, 16, 10, 5, 0.0, None, None)
    assert rc == -2
    # a generic basis without its device copy is refused before anything is launched
    from aas_enhancement_b200 import _lib
    dense = MelPlan(np.ones((40, 161), dtype=np.float32))
    io = _lib.make_io(flags=5, n=1, tmax=10, wave=16, lengths=16, mask_r=16, mask_i=16, mask_stride_n=1610,
                      mask_stride_f=10, window=16, out=16, stats=16)
    assert lib.aas_lmfb_forward_ex(dense.handle, ctypes.byref(io)) == -5
    io.struct_size = 8
    assert lib.aas_lmfb_forward_ex(plan.handle, ctypes.byref(io)) == -3


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


def test_build_reads_the_stack_frames_of_the_hot_kernels():
    """aas_enhancement_b200/build.py keeps a build only if ptxas reports no stack frame for the kernels every
    default call launches (a kernel with local memory costs launch time and spills in the tile loop)."""
    from aas_enhancement_b200 import build as b
    log = "\n".join([
        "ptxas info    : Compiling entry function '_ZN8aas_lmfb7lmfb_k1ILi1ELb0ELi5ELi3ELb0ELb0EEEvNS_6K1ArgsENS_5TabOfIXT0_EE5ParamE' for 'sm_100a'",
        "ptxas info    : Function properties for _ZN8aas_lmfb7lmfb_k1ILi1ELb0ELi5ELi3ELb0ELb0EEEvNS_6K1ArgsENS_5TabOfIXT0_EE5ParamE",
        "    16 bytes stack frame, 12 bytes spill stores, 16 bytes spill loads",
        "ptxas info    : Used 128 registers, used 1 barriers, 16 bytes cumulative stack size",
        "ptxas info    : Compiling entry function '_ZN8aas_lmfb7lmfb_k1ILi1ELb1ELi8ELi2ELb0ELb0EEEvNS_6K1ArgsENS_5TabOfIXT0_EE5ParamE' for 'sm_100a'",
        "    8 bytes stack frame, 8 bytes spill stores, 8 bytes spill loads",
        "ptxas info    : Used 128 registers, used 1 barriers",
        "ptxas info    : Compiling entry function '_ZN8aas_lmfb7lmfb_k1ILi1ELb1ELi4ELi3ELb1ELb0EEEvNS_6K1ArgsENS_5TabOfIXT0_EE5ParamE' for 'sm_100a'",
        "    64 bytes stack frame, 60 bytes spill stores, 60 bytes spill loads",
        "ptxas info    : Used 168 registers, used 1 barriers",
    ])
    rep = b.stack_report(log)
    assert len(rep) == 3
    assert [v for k, v in rep.items() if "ELi5ELi3ELb0ELb0E" in k] == [[16, 128]]
    # five-warp kernel: 16 bytes; eight-warp kernel: 8 bytes, weighted a thousandfold; the wave-gradient kernel is not a hot one
    assert b.hot_stack_bytes(rep) == 16 + 8 * 1000
    assert b.hot_stack_bytes(b.stack_report(log.replace("    8 bytes stack", "    0 bytes stack"))) <= b.GOOD_ENOUGH
    assert b.SPLITS[0] == 1                      # the reproducible single-module build is tried first

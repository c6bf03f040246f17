"""ctypes binding of libaas_lmfb.so (the C ABI declared in include/aas_lmfb.h).

There is deliberately no fallback: if the CUDA library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libaas_lmfb.so")

ABI_VERSION = 2
N_FFT, HOP, N_BINS = 320, 160, 161

MASK_MODES = {"none": 0, "reim": 1, "power": 2}
CMVN_MODES = {"none": 0 << 2, "per_bin": 1 << 2, "global": 2 << 2}
WAVE_I16 = 1 << 4

EXPORTS = ("aas_lmfb_abi_version", "aas_lmfb_strerror", "aas_lmfb_plan_create",
           "aas_lmfb_plan_destroy", "aas_lmfb_plan_info", "aas_lmfb_plan_set_tuning",
           "aas_lmfb_plan_tables_bytes", "aas_lmfb_plan_upload",
           "aas_lmfb_workspace_bytes", "aas_lmfb_forward_ex", "aas_lmfb_backward_ex", "aas_lmfb_forward",
           "aas_lmfb_backward", "aas_lmfb_backward_wave", "aas_lmfb_stft", "aas_l1_partial_count", "aas_l1_abs_sum", "aas_l1_rows_sum", "aas_l1_abs_grad")

_lock = threading.Lock()
_lib = None

_vp, _i32, _i64, _u32, _f32 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32,
                               ctypes.c_float)


class IO(ctypes.Structure):
    """``aas_lmfb_io`` of include/aas_lmfb.h, field for field."""
    _fields_ = [("struct_size", _u32), ("flags", _u32), ("device", _i32), ("n", _i32), ("n_ch", _i32),
                ("tmax", _i32), ("eps", _f32), ("reserved_", _i32),
                ("wave", _vp), ("wave_stride", _i64), ("wave_stride_ch", _i64), ("wave_len", _i64),
                ("lengths", _vp), ("mask_r", _vp), ("mask_i", _vp), ("mask_stride_n", _i64),
                ("mask_stride_f", _i64), ("window", _vp), ("mel_dev", _vp), ("out", _vp), ("stats", _vp),
                ("grad_out", _vp), ("grad_mask_r", _vp), ("grad_mask_i", _vp), ("grad_wave", _vp),
                ("workspace", _vp), ("cuda_stream", _vp), ("prof", _vp), ("tables", _vp),
                ("l1_target", _vp), ("l1_rows", _vp), ("frame_lens", _vp)]


def make_io(**kw) -> IO:
    io = IO()
    io.struct_size = ctypes.sizeof(IO)
    io.device = -1
    io.n_ch = 1
    for k, v in kw.items():
        setattr(io, k, v)
    return io


def set_library_path(path: str) -> None:
    """Development A/B runs only (bench.py --lib): load another build of the library.  Must be
    called before the first :func:`load`."""
    global LIB_PATH
    if _lib is not None:
        raise RuntimeError("the library is already loaded")
    LIB_PATH = path


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m aas_enhancement_b200.build` "
                "(the LMFB front-end is CUDA-only; there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.aas_lmfb_abi_version.restype = _i32
        lib.aas_lmfb_strerror.restype = ctypes.c_char_p
        lib.aas_lmfb_strerror.argtypes = [_i32]
        lib.aas_lmfb_plan_create.restype = _vp
        lib.aas_lmfb_plan_create.argtypes = [_vp, _i32, _i32, ctypes.POINTER(_i32)]
        lib.aas_lmfb_plan_destroy.restype = None
        lib.aas_lmfb_plan_destroy.argtypes = [_vp]
        lib.aas_lmfb_plan_info.restype = _i32
        lib.aas_lmfb_plan_info.argtypes = [_vp, ctypes.POINTER(_i32), ctypes.POINTER(_i32), ctypes.POINTER(_i32)]
        lib.aas_lmfb_plan_set_tuning.restype = _i32
        lib.aas_lmfb_plan_set_tuning.argtypes = [_vp, _i32, _i32, _i32]
        lib.aas_lmfb_plan_tables_bytes.restype = ctypes.c_size_t
        lib.aas_lmfb_plan_tables_bytes.argtypes = [_vp]
        lib.aas_lmfb_plan_upload.restype = _i32
        lib.aas_lmfb_plan_upload.argtypes = [_vp, _vp, _vp]
        lib.aas_lmfb_workspace_bytes.restype = ctypes.c_size_t
        lib.aas_lmfb_workspace_bytes.argtypes = [_vp, _i32, _i32, _u32]
        lib.aas_lmfb_forward_ex.restype = _i32
        lib.aas_lmfb_forward_ex.argtypes = [_vp, ctypes.POINTER(IO)]
        lib.aas_lmfb_backward_ex.restype = _i32
        lib.aas_lmfb_backward_ex.argtypes = [_vp, ctypes.POINTER(IO)]
        lib.aas_lmfb_forward.restype = _i32
        lib.aas_lmfb_forward.argtypes = [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _i64, _vp,
                                         _vp, _vp, _i32, _u32, _f32, _vp, _vp]
        lib.aas_lmfb_backward.restype = _i32
        lib.aas_lmfb_backward.argtypes = [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _i64, _vp,
                                          _vp, _vp, _vp, _vp, _vp, _vp, _i32, _u32, _f32, _vp, _vp]
        lib.aas_lmfb_backward_wave.restype = _i32
        lib.aas_lmfb_backward_wave.argtypes = [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _i64, _vp,
                                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _u32, _f32, _vp]
        lib.aas_lmfb_stft.restype = _i32
        lib.aas_lmfb_stft.argtypes = [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _i32, _vp]
        lib.aas_l1_partial_count.restype = _i32
        lib.aas_l1_abs_sum.restype = _i32
        lib.aas_l1_abs_sum.argtypes = [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]
        lib.aas_l1_rows_sum.restype = _i32
        lib.aas_l1_rows_sum.argtypes = [_vp, _i32, _vp, _vp]
        lib.aas_l1_abs_grad.restype = _i32
        lib.aas_l1_abs_grad.argtypes = [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]
        if lib.aas_lmfb_abi_version() != ABI_VERSION:
            raise RuntimeError("libaas_lmfb.so ABI version mismatch; rebuild it")
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(load().aas_lmfb_strerror(rc).decode() + f" (code {rc})")

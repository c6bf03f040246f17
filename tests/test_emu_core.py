"""CPU emulation of the device FFT code (lane = frame, PFA 5x32 + real split) against the
float64 oracle.  Needs only g++; no GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import lmfb_oracle as orc
import _synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "emu_core.cpp")
INC = os.path.join(ROOT, "aas_enhancement_b200", "csrc")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu_core.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-I", INC, SRC, "-o", so])
    lib = ctypes.CDLL(so)
    vp = ctypes.c_void_p
    lib.emu_k1.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_longlong,
                           vp, vp, ctypes.c_longlong, ctypes.c_longlong, vp, vp, ctypes.c_int, vp, vp, vp, vp,
                           ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_longlong]
    lib.emu_mel_tables.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
    return lib


def _k1(lib, warps, bwd, mode, b, mel, window, dE=None, vec_ok=1, want_wave_grad=False):
    """wave (N, L) or (N, nCH, L); masks (N, nCH*161, T)."""
    modes = {"none": 0, "reim": 1, "power": 2}
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    wave, mr, mi = f32(b["wave"]), f32(b["mask_r"]), f32(b["mask_i"])
    lens = np.ascontiguousarray(b["lengths"], dtype=np.int32)
    n, tmax = wave.shape[0], b["tmax"]
    n_ch = wave.shape[1] if wave.ndim == 3 else 1
    melf, win = f32(mel), f32(window)
    out = np.full((n, mel.shape[0], tmax), np.nan, dtype=np.float32)
    gr = np.full_like(mr, np.nan)
    gi = np.full_like(mi, np.nan)
    dE = f32(dE) if dE is not None else np.zeros_like(out)
    gw = np.zeros_like(wave) if want_wave_grad else None
    rc = lib.emu_k1(warps, int(bwd), modes[mode], wave.ctypes.data, lens.ctypes.data, n, n_ch * wave.shape[-1],
                    mr.ctypes.data if mode != "none" else None,
                    mi.ctypes.data if mode == "reim" else None,
                    mr.shape[1] * mr.shape[2], mr.shape[2], win.ctypes.data, melf.ctypes.data,
                    mel.shape[0], out.ctypes.data, dE.ctypes.data,
                    gr.ctypes.data if mode != "none" else None, gi.ctypes.data if mode == "reim" else None,
                    tmax, vec_ok, gw.ctypes.data if want_wave_grad else None, n_ch, wave.shape[-1])
    assert rc == 0
    if want_wave_grad:
        return out, gr, gi, gw
    return out, gr, gi


@pytest.mark.parametrize("length,sym", [(16000, True), (5000, True), (4999, False), (333, True),
                                        (161, True), (100, True), (1, True)])
def test_emulated_stft_matches_oracle(emu, length, sym):
    """Unit mel basis picks single bins: with masks (1,0) / (0,1) the forward output is
    log1p(Re^2) / log1p(Im^2) of every bin, i.e. the STFT itself (edge tiles, unaligned path)."""
    rs = np.random.RandomState(length)
    wave = (0.1 * rs.randn(1, length + 7)).astype(np.float32)   # trailing garbage must not be read
    window = orc.hamming_window(320, sym)
    spec = orc.stft_frames(wave[0], length, window.astype(np.float32).astype(np.float64))
    t_i = spec.shape[1]
    scale = np.abs(spec).max() ** 2
    # 80 single-bin "filters" per run (bins lo..lo+79), two runs cover all 161 bins
    for lo in (0, 81):
        nb = min(80, 161 - lo)
        mel = np.zeros((nb, 161))
        mel[np.arange(nb), lo + np.arange(nb)] = 1.0
        for which, vec_ok in (("re", 1), ("im", 0)):
            b = dict(wave=wave, lengths=np.array([length]), tmax=t_i,
                     mask_r=np.full((1, 161, t_i), 1.0 if which == "re" else 0.0, np.float32),
                     mask_i=np.full((1, 161, t_i), 0.0 if which == "re" else 1.0, np.float32))
            y, _, _ = _k1(emu, 4, 0, "reim", b, mel, window, vec_ok=vec_ok)
            got = np.expm1(y[0].astype(np.float64))
            ref = (spec.real if which == "re" else spec.imag)[lo:lo + nb] ** 2
            assert np.abs(got - ref).max() / scale < 3e-6, (lo, which)


@pytest.mark.parametrize("warps", [1, 2, 3, 4, 5, 6, 8])
@pytest.mark.parametrize("mode", ["reim", "power", "none"])
def test_emulated_k1_forward_and_backward(emu, mode, warps):
    b = _synth.make_batch(3, 5000, seed=17, ragged=True, tonal=(mode == "power"))
    mel, window = orc.mel_filterbank(), orc.hamming_window()
    mel32 = mel.astype(np.float32).astype(np.float64)
    win32 = window.astype(np.float32).astype(np.float64)
    mr = b["mask_r"] if mode != "none" else None
    mi = b["mask_i"] if mode == "reim" else None
    y_ref, fl = orc.lmfb_forward(b["wave"], b["lengths"], mr, mi, mel32, win32, mask_mode=mode,
                                 cmvn_mode="none")
    y, _, _ = _k1(emu, warps, 0, mode, b, mel, window)
    assert not np.isnan(y).any()
    assert orc.rel_err(y, y_ref) < 2e-5
    for i in range(3):
        assert np.all(y[i, :, fl[i]:] == 0.0)
    if mode == "none":
        return
    g_ref = orc.lmfb_grads(b["wave"], b["lengths"], mr, mi, b["grad_out"], mel32, win32,
                           mask_mode=mode, cmvn_mode="none")
    dE = b["grad_out"].astype(np.float64) * np.exp(-y_ref)
    for i in range(3):
        dE[i, :, fl[i]:] = 0.0
    _, gr, gi = _k1(emu, warps, 1, mode, b, mel, window, dE=dE)
    assert not np.isnan(gr).any()
    assert orc.rel_err(gr, g_ref["grad_mask_r"]) < 2e-5
    if mode == "reim":
        assert not np.isnan(gi).any()
        assert orc.rel_err(gi, g_ref["grad_mask_i"]) < 2e-5


@pytest.mark.parametrize("n_mels,warps", [(80, 3), (64, 5), (2, 4)])
def test_emulated_forward_other_bases(emu, n_mels, warps):
    """Bases in which several filters end on the same bin (hand-over of more than one filter per
    bin, the rolled phase-3 path) and the two-filter minimum."""
    b = _synth.make_batch(2, 4000, seed=n_mels, ragged=True)
    mel = orc.mel_filterbank(n_mels=n_mels) if n_mels > 2 else np.stack([np.linspace(1, 0, 161), np.linspace(0, 1, 161)])
    window = orc.hamming_window()
    mel32 = mel.astype(np.float32).astype(np.float64)
    win32 = window.astype(np.float32).astype(np.float64)
    y_ref, fl = orc.lmfb_forward(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], mel32, win32,
                                 mask_mode="reim", cmvn_mode="none")
    y, _, _ = _k1(emu, warps, 0, "reim", b, mel, window)
    assert not np.isnan(y).any()
    assert orc.rel_err(y, y_ref) < 2e-5
    g = np.random.RandomState(3).randn(*y_ref.shape)
    g_ref = orc.lmfb_grads(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], g, mel32, win32,
                           mask_mode="reim", cmvn_mode="none")
    dE = g * np.exp(-y_ref)
    for i in range(2):
        dE[i, :, fl[i]:] = 0.0
    _, gr, gi = _k1(emu, warps, 1, "reim", b, mel, window, dE=dE)
    assert orc.rel_err(gr, g_ref["grad_mask_r"]) < 2e-5
    assert orc.rel_err(gi, g_ref["grad_mask_i"]) < 2e-5


@pytest.mark.parametrize("mode,warps", [("none", 3), ("reim", 4), ("power", 2), ("reim", 5)])
@pytest.mark.parametrize("length", [5000, 4999, 161, 700])
def test_emulated_wave_gradient(emu, mode, warps, length):
    """Gradient into the waveform (adjoint split / DFT5 / FFT32 / window / overlap-add with the
    reflect padding folded back) against float64 autograd; the mask gradients must not change."""
    b = _synth.make_batch(2, length, seed=length + warps, ragged=True)
    mel, window = orc.mel_filterbank(), orc.hamming_window()
    mel32 = mel.astype(np.float32).astype(np.float64)
    win32 = window.astype(np.float32).astype(np.float64)
    mr = b["mask_r"] if mode != "none" else None
    mi = b["mask_i"] if mode == "reim" else None
    y_ref, fl = orc.lmfb_forward(b["wave"], b["lengths"], mr, mi, mel32, win32, mask_mode=mode, cmvn_mode="none")
    g_ref = orc.lmfb_grads(b["wave"], b["lengths"], mr, mi, b["grad_out"], mel32, win32,
                           mask_mode=mode, cmvn_mode="none", want_wave_grad=True)
    dE = b["grad_out"].astype(np.float64) * np.exp(-y_ref)
    for i in range(2):
        dE[i, :, fl[i]:] = 0.0
    _, gr, gi, gw = _k1(emu, warps, 1, mode, b, mel, window, dE=dE, want_wave_grad=True, vec_ok=int(length % 2 == 0))
    ref = g_ref["grad_wave"]
    for i in range(2):
        li = int(b["lengths"][i])
        assert np.all(gw[i, li:] == 0.0)                       # nothing beyond the utterance
    assert orc.rel_err(gw, ref) < 5e-5
    tol = 2e-5 if length > 320 else 2e-4      # two reflect-padded frames: the imaginary parts are cancellation noise
    if mode != "none":
        assert orc.rel_err(gr, g_ref["grad_mask_r"]) < tol
    if mode == "reim":
        assert orc.rel_err(gi, g_ref["grad_mask_i"]) < tol


def _tables(emu, mel, warps=4):
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    fwd = np.zeros_like(mel); bwd = np.zeros_like(mel)
    walkable = ctypes.c_int(-1); banded = ctypes.c_int(-1)
    lohi = np.zeros(16, np.int32)
    rc = emu.emu_mel_tables(mel.ctypes.data, mel.shape[0], warps, fwd.ctypes.data, bwd.ctypes.data,
                            ctypes.byref(walkable), ctypes.byref(banded), lohi.ctypes.data)
    assert rc == 0
    return fwd, bwd, walkable.value, banded.value


def test_phase3_filter_partition(emu):
    """Forward phase 3: the filters are dealt to the warps as contiguous runs that cover every filter exactly
    once (mel_band.hpp: set_warp_ranges), whatever the basis and the number of warps."""
    for n_mels in (40, 23, 64, 80, 128, 2):
        mel = np.ascontiguousarray(orc.mel_filterbank(n_mels=n_mels) if n_mels > 2 else
                                   np.stack([np.linspace(1, 0, 161), np.linspace(0, 1, 161)]), dtype=np.float32)
        for warps in (1, 2, 3, 4, 5, 6, 8):
            fwd = np.zeros_like(mel); bwd = np.zeros_like(mel)
            walkable = ctypes.c_int(-1); banded = ctypes.c_int(-1)
            lohi = np.zeros(16, np.int32)
            assert emu.emu_mel_tables(mel.ctypes.data, n_mels, warps, fwd.ctypes.data, bwd.ctypes.data,
                                      ctypes.byref(walkable), ctypes.byref(banded), lohi.ctypes.data) == 0
            owned = []
            for w in range(8):
                lo, hi = int(lohi[2 * w]), int(lohi[2 * w + 1])
                if w >= warps:
                    assert lo > hi                       # warps that do not exist own nothing
                owned += list(range(lo, hi + 1))
            assert owned == list(range(n_mels)), (n_mels, warps, lohi)


def test_mel_tables_reproduce_the_basis(emu):
    """The forward tables (walk or rows) and the backward (d, d+1) table, expanded back into dense
    matrices, are the basis itself: triangular filterbanks are walkable + banded, anything else is generic."""
    for n_mels in (40, 23, 64, 80, 2):
        for warps in (1, 2, 4, 5):
            mel = np.ascontiguousarray(orc.mel_filterbank(n_mels=n_mels) if n_mels > 2 else
                                       np.stack([np.linspace(1, 0, 161), np.linspace(0, 1, 161)]), dtype=np.float32)
            fwd, bwd, walkable, banded = _tables(emu, mel, warps)
            assert walkable == 1 and np.array_equal(fwd, mel)
            assert banded == 1 and np.array_equal(bwd, mel)
    dense = np.random.RandomState(0).rand(40, 161).astype(np.float32)
    fwd, _, walkable, banded = _tables(emu, dense)
    assert walkable == 0 and banded == 0 and np.array_equal(fwd, dense)
    perm = orc.mel_filterbank()[np.random.RandomState(1).permutation(40)]   # re-ordered filters
    fwd, _, walkable, banded = _tables(emu, perm)
    assert walkable == 0 and banded == 0 and np.array_equal(fwd, perm.astype(np.float32))


@pytest.mark.parametrize("kind,warps", [("dense", 3), ("permuted", 4), ("ones", 2), ("gappy", 5)])
def test_emulated_generic_bases(emu, kind, warps):
    """Any (M, 161) matrix is a valid basis (the reference applies it with a k=1 conv1d, model.py:196):
    dense, re-ordered, all-ones (MelPlan(np.ones((40, 161)))), and rows with holes."""
    rs = np.random.RandomState(11)
    if kind == "dense":
        mel = rs.rand(40, 161) * 0.02
    elif kind == "permuted":
        mel = orc.mel_filterbank()[rs.permutation(40)]
    elif kind == "ones":
        mel = np.ones((40, 161)) * 0.01
    else:
        mel = orc.mel_filterbank() * (rs.rand(40, 161) > 0.3)
        mel[7] = 0.0                                                     # an empty filter
    b = _synth.make_batch(2, 4000, seed=5, ragged=True)
    window = orc.hamming_window()
    mel32 = mel.astype(np.float32).astype(np.float64)
    win32 = window.astype(np.float32).astype(np.float64)
    y_ref, fl = orc.lmfb_forward(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], mel32, win32,
                                 mask_mode="reim", cmvn_mode="none")
    y, _, _ = _k1(emu, warps, 0, "reim", b, mel, window)
    assert not np.isnan(y).any()
    assert orc.rel_err(y, y_ref) < 2e-5
    g_ref = orc.lmfb_grads(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], b["grad_out"], mel32, win32,
                           mask_mode="reim", cmvn_mode="none")
    dE = b["grad_out"].astype(np.float64) * np.exp(-y_ref)
    for i in range(2):
        dE[i, :, fl[i]:] = 0.0
    _, gr, gi = _k1(emu, warps, 1, "reim", b, mel, window, dE=dE)
    assert orc.rel_err(gr, g_ref["grad_mask_r"]) < 2e-5
    assert orc.rel_err(gi, g_ref["grad_mask_i"]) < 2e-5


@pytest.mark.parametrize("n_ch,mode,warps", [(2, "reim", 3), (3, "power", 4), (2, "none", 2)])
def test_emulated_multi_channel(emu, n_ch, mode, warps):
    """BRNNmultiCH with nCH > 1 (model.py:160-167, :186-198): masks (N, nCH*F, T), the basis repeats over
    the channels, i.e. the masked powers of the channels are summed before the mel projection."""
    rs = np.random.RandomState(n_ch)
    base = _synth.make_batch(2, 3000, seed=31 + n_ch, ragged=True)
    n, tmax, L = 2, base["tmax"], base["wave"].shape[1]
    wave = np.zeros((n, n_ch, L), np.float32)
    for i in range(n):
        li = int(base["lengths"][i])
        wave[i, :, :li] = (0.1 * rs.randn(n_ch, li)).astype(np.float32)
    b = dict(wave=wave, lengths=base["lengths"], tmax=tmax,
             mask_r=rs.uniform(0, 1, (n, n_ch * 161, tmax)).astype(np.float32),
             mask_i=rs.uniform(0, 1, (n, n_ch * 161, tmax)).astype(np.float32))
    mel, window = orc.mel_filterbank(), orc.hamming_window()
    mel32 = mel.astype(np.float32).astype(np.float64)
    win32 = window.astype(np.float32).astype(np.float64)
    mr = b["mask_r"] if mode != "none" else None
    mi = b["mask_i"] if mode == "reim" else None
    y_ref, fl = orc.lmfb_forward(wave, b["lengths"], mr, mi, mel32, win32, mask_mode=mode, cmvn_mode="none")
    y, _, _ = _k1(emu, warps, 0, mode, b, mel, window)
    assert orc.rel_err(y, y_ref) < 2e-5
    if mode == "none":
        return
    g = rs.randn(n, 40, tmax)
    g_ref = orc.lmfb_grads(wave, b["lengths"], mr, mi, g, mel32, win32, mask_mode=mode, cmvn_mode="none")
    dE = g * np.exp(-y_ref)
    for i in range(n):
        dE[i, :, fl[i]:] = 0.0
    _, gr, gi = _k1(emu, warps, 1, mode, b, mel, window, dE=dE)
    assert orc.rel_err(gr, g_ref["grad_mask_r"]) < 2e-5
    if mode == "reim":
        assert orc.rel_err(gi, g_ref["grad_mask_i"]) < 2e-5

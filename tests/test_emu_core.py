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
                           ctypes.c_int, ctypes.c_int, vp]
    lib.emu_mel_band.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    return lib


def _k1(lib, warps, bwd, mode, b, mel, window, dE=None, vec_ok=1, want_wave_grad=False):
    modes = {"none": 0, "reim": 1, "power": 2}
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    wave, mr, mi = f32(b["wave"]), f32(b["mask_r"]), f32(b["mask_i"])
    lens = np.ascontiguousarray(b["lengths"], dtype=np.int32)
    n, tmax = wave.shape[0], b["tmax"]
    melf, win = f32(mel), f32(window)
    out = np.full((n, mel.shape[0], tmax), np.nan, dtype=np.float32)
    gr = np.full_like(mr, np.nan)
    gi = np.full_like(mi, np.nan)
    dE = f32(dE) if dE is not None else np.zeros_like(out)
    gw = np.zeros_like(wave) if want_wave_grad else None
    rc = lib.emu_k1(warps, int(bwd), modes[mode], wave.ctypes.data, lens.ctypes.data, n, wave.shape[1],
                    mr.ctypes.data if mode != "none" else None,
                    mi.ctypes.data if mode == "reim" else None,
                    mr.shape[1] * mr.shape[2], mr.shape[2], win.ctypes.data, melf.ctypes.data,
                    mel.shape[0], out.ctypes.data, dE.ctypes.data,
                    gr.ctypes.data if mode != "none" else None, gi.ctypes.data if mode == "reim" else None,
                    tmax, vec_ok, gw.ctypes.data if want_wave_grad else None)
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


@pytest.mark.parametrize("warps", [1, 2, 3, 4, 5])
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


def test_mel_band_tables(emu):
    for n_mels in (40, 23, 64, 80):
        for warps in (1, 2, 4, 5):
            mel = np.ascontiguousarray(orc.mel_filterbank(n_mels=n_mels), dtype=np.float32)
            wl = np.zeros(161, np.float32); wh = np.zeros(161, np.float32)
            ml = np.zeros(161, np.int32); lohi = np.zeros(16, np.int32)
            assert emu.emu_mel_band(mel.ctypes.data, n_mels, warps, wl.ctypes.data, wh.ctypes.data,
                                    ml.ctypes.data, lohi.ctypes.data) == 0
            rebuilt = np.zeros_like(mel)
            for f in range(161):
                if ml[f] < n_mels:
                    rebuilt[ml[f], f] += 4 * wl[f]
                if ml[f] + 1 < n_mels:
                    rebuilt[ml[f] + 1, f] += 4 * wh[f]
            assert np.array_equal(rebuilt, mel)
            assert np.all(np.diff(ml) >= 0)
    dense = np.ones((40, 161), dtype=np.float32)
    wl = np.zeros(161, np.float32); wh = np.zeros(161, np.float32)
    ml = np.zeros(161, np.int32); lohi = np.zeros(16, np.int32)
    assert emu.emu_mel_band(dense.ctypes.data, 40, 4, wl.ctypes.data, wh.ctypes.data, ml.ctypes.data,
                            lohi.ctypes.data) == -1

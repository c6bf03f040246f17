"""Drop-in differentiable LMFB front-end (host side of the C ABI).

Mirrors the operator surface the reference exposes for this path:

* ``BRNNmultiCH(..., mel_basis)`` keeps the mel basis as a module buffer and its ``forward``
  ends in mask -> power -> mel -> log1p on ``(N, F, T)`` tensors
  (Speech_enhancement_by_AAS/model.py:148, :167, :186-198)  ->  :class:`LMFBFrontEnd`.
* trainers consume ``(N, 40, T)`` fp32, channel-first, time-last
  (trainer_AAS.py:136-142, loader_functions.py:56)           ->  the output layout here.
* front-end parameters 16 kHz / 20 ms Hamming / 10 ms stride / 40 mels
  (AM_training/train.py:39-42, :56)                           ->  defaults below.

All arithmetic runs in the hand-written sm_100a kernels behind ``libaas_lmfb.so``; torch is
used for device memory and streams only.  There is no CPU path.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _lib

N_FFT, HOP, N_BINS = _lib.N_FFT, _lib.HOP, _lib.N_BINS
SAMPLE_RATE = 16000


# --------------------------------------------------------------------------- parameters
def hamming_window(n: int = N_FFT, sym: bool = True) -> np.ndarray:
    """'hamming' is pinned by AM_training/train.py:42; symmetric is what deepspeech.pytorch
    hands to librosa (scipy.signal.hamming).  float64."""
    k = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / ((n - 1) if sym else n))


def slaney_mel_basis(sr: int = SAMPLE_RATE, n_fft: int = N_FFT, n_mels: int = 40,
                     fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """Default mel basis: Slaney scale, area-normalised triangles (librosa.filters.mel
    defaults), (n_mels, n_fft//2+1) float64.  The reference takes the basis as a constructor
    argument (model.py:148); pass your own to :class:`LMFBFrontEnd` to override."""
    fmax = sr / 2.0 if fmax is None else fmax
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz2mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def mel2hz(m):
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    n_bins = n_fft // 2 + 1
    freqs = np.linspace(0.0, sr / 2.0, n_bins)
    hz = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(hz)
    ramps = hz[:, None] - freqs[None, :]
    w = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        w[i] = np.maximum(0.0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    return w * (2.0 / (hz[2:] - hz[:-2]))[:, None]


class MelPlan:
    """Host-side plan for one mel basis (wraps ``aas_lmfb_plan_create``).  Keeps the host matrix,
    so that copies (``copy.deepcopy``, pickling, ``torch.save`` of a whole module) rebuild the
    native handle instead of trying to serialise it."""

    def __init__(self, mel_basis):
        mel = np.ascontiguousarray(
            mel_basis.detach().cpu().numpy() if isinstance(mel_basis, torch.Tensor) else mel_basis,
            dtype=np.float32)
        if mel.ndim != 2:
            raise ValueError("mel_basis must be (n_mels, 161)")
        lib = _lib.load()
        status = ctypes.c_int(0)
        self._lib = lib
        self.mel = mel
        self._tables = {}
        self.n_mels = int(mel.shape[0])
        self.handle = lib.aas_lmfb_plan_create(mel.ctypes.data, mel.shape[0], mel.shape[1],
                                               ctypes.byref(status))
        if not self.handle:
            _lib.check(status.value)
            raise RuntimeError("aas_lmfb_plan_create failed")

    def tables(self, device):
        """Device-resident lookup tables of this plan on ``device`` (uploaded on first use; the
        kernels fetch them with one bulk copy per thread block).  Returns the device pointer."""
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        t = self._tables.get(key)
        if t is None:
            nbytes = self._lib.aas_lmfb_plan_tables_bytes(self.handle)
            t = torch.empty(nbytes, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                _lib.check(self._lib.aas_lmfb_plan_upload(self.handle, t.data_ptr(),
                                                          torch.cuda.current_stream(device).cuda_stream))
                torch.cuda.current_stream(device).synchronize()     # pageable host source: do not race the plan's lifetime
            self._tables[key] = t
        return t.data_ptr()

    def __reduce__(self):
        return (MelPlan, (self.mel,))

    def __deepcopy__(self, memo):
        return MelPlan(self.mel)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.aas_lmfb_plan_destroy(h)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _rows_disjoint(m):
    """(N, 161, T) tensor whose rows do not overlap in memory (gradients are written with the same
    strides, so an expanded / overlapping view would make utterances race on the same words)."""
    n, f, t = m.shape
    return m.stride(1) >= t and (n == 1 or m.stride(0) >= f * m.stride(1))


def _check_f32_cuda(name, t, dev):
    if t.dtype != torch.float32 or t.device != dev:
        raise TypeError(f"{name} must be a float32 tensor on {dev}")


class LMFB(torch.autograd.Function):
    """``Z, frame_lens = LMFB.apply(wave, lengths, mask_r, mask_i, plan, window, mask_mode,
    cmvn_mode, eps, tmax, mel_dev)``

    wave (N, Lmax) f32 cuda zero-padded (or int16 PCM, value / 32768: converted inside the kernel, half
    the bytes on the wire) -- or (N, nCH, Lmax) for multi-channel input, with masks
    (N, nCH*161, Tmax) as in ``BRNNmultiCH`` (model.py:160-167, :186-198: the basis repeats over the
    channels, i.e. the masked powers are summed); lengths (N,) int32 cuda (samples); masks
    (N, 161, Tmax) f32 or None; plan a :class:`MelPlan`; window (320,) f32 cuda; ``mel_dev`` the
    (M, 161) basis on the device (only read for bases off the fast path).
    Returns Z (N, M, Tmax) f32 with frames ``t >= T_i`` exactly zero, and frame_lens (N,)
    int32 (``T_i = 1 + L_i // 160``).  Gradients flow to mask_r / mask_i and, when ``wave``
    requires grad, to the waveform.
    """

    @staticmethod
    def forward(ctx, wave, lengths, mask_r, mask_i, plan, window, mask_mode="reim",
                cmvn_mode="per_bin", eps=0.0, tmax=None, mel_dev=None, l1_target=None):
        if not wave.is_cuda:
            raise RuntimeError("LMFB is CUDA-only (sm_100a); there is no CPU fallback")
        dev = wave.device
        lib = _lib.load()
        if mask_mode not in _lib.MASK_MODES or cmvn_mode not in _lib.CMVN_MODES:
            raise ValueError(f"bad mask_mode/cmvn_mode: {mask_mode!r}/{cmvn_mode!r}")
        if wave.dtype == torch.int16:                  # PCM as stored in wave files: converted inside the kernel
            if wave.device != dev:
                raise TypeError("wave must live on one device")
        else:
            _check_f32_cuda("wave", wave, dev)
        _check_f32_cuda("window", window, dev)
        if wave.dim() not in (2, 3) or wave.stride(-1) != 1:
            raise ValueError("wave must be (N, Lmax) or (N, nCH, Lmax) with unit sample stride")
        n = wave.shape[0]
        n_ch = wave.shape[1] if wave.dim() == 3 else 1
        lmax = wave.shape[-1]
        if (n > 1 and wave.stride(0) < n_ch * lmax) or (wave.dim() == 3 and n_ch > 1 and wave.stride(1) < lmax):
            wave = wave.contiguous()                   # overlapping rows (e.g. an expanded view)
        if lengths.dtype != torch.int32 or lengths.device != dev:
            lengths = lengths.to(device=dev, dtype=torch.int32)
        lengths = lengths.contiguous()
        window = window.contiguous()
        rows = n_ch * N_BINS
        use_r = mask_mode in ("reim", "power")
        use_i = mask_mode == "reim"
        if (use_r and mask_r is None) or (use_i and mask_i is None):
            raise ValueError(f"mask_mode={mask_mode!r} needs its mask tensor(s)")
        mask_r = mask_r if use_r else None
        mask_i = mask_i if use_i else None
        msn = msf = 0
        if use_r:
            _check_f32_cuda("mask_r", mask_r, dev)
            if mask_r.dim() != 3 or mask_r.shape[0] != n or mask_r.shape[1] != rows:
                raise ValueError("mask_r must be (N, nCH*161, Tmax)")
            if mask_r.stride(2) != 1 or not _rows_disjoint(mask_r):
                mask_r = mask_r.contiguous()           # e.g. a mask expanded over the batch (stride 0)
            if tmax is None:
                tmax = mask_r.shape[2]
            elif mask_r.shape[2] != tmax:
                raise ValueError("mask_r.shape[2] != tmax")
            msn, msf = mask_r.stride(0), mask_r.stride(1)
        if use_i:
            _check_f32_cuda("mask_i", mask_i, dev)
            if mask_i.shape != mask_r.shape:
                raise ValueError("mask_i must have the shape of mask_r")
            if mask_i.stride() != mask_r.stride() or not _rows_disjoint(mask_i):
                mask_i = mask_i.contiguous()
                mask_r = mask_r.contiguous()
                msn, msf = mask_r.stride(0), mask_r.stride(1)
        if tmax is None:
            tmax = 1 + lmax // HOP
        tmax = int(tmax)
        if mel_dev is not None:
            _check_f32_cuda("mel_dev", mel_dev, dev)
            mel_dev = mel_dev.contiguous()
        flags = _lib.MASK_MODES[mask_mode] | _lib.CMVN_MODES[cmvn_mode] | (_lib.WAVE_I16 if wave.dtype == torch.int16 else 0)
        out = torch.empty((n, plan.n_mels, tmax), dtype=torch.float32, device=dev)
        stats = torch.empty((n, plan.n_mels, 2), dtype=torch.float32, device=dev)
        frame_lens = torch.empty((n,), dtype=torch.int32, device=dev)      # written by the kernel (no torch arithmetic per call)
        l1_rows = None
        if l1_target is not None:                      # L1Loss_mask epilogue of the CMVN kernel: Z is read once
            if cmvn_mode == "none":
                raise ValueError("l1_target needs a CMVN mode (the epilogue lives in the CMVN kernel)")
            _check_f32_cuda("l1_target", l1_target, dev)
            if tuple(l1_target.shape) != (n, plan.n_mels, tmax):
                raise ValueError("l1_target must be (N, n_mels, Tmax)")
            l1_target = l1_target.detach().contiguous()
            l1_rows = torch.empty((n, plan.n_mels), dtype=torch.float32, device=dev)
        io = _lib.make_io(flags=flags, device=dev.index if dev.index is not None else -1, n=n, n_ch=n_ch, tmax=tmax,
                          eps=float(eps), wave=wave.data_ptr(), wave_stride=_row_stride(wave, n_ch),
                          wave_stride_ch=wave.stride(1) if wave.dim() == 3 else 0, wave_len=lmax,
                          lengths=lengths.data_ptr(), mask_r=_ptr(mask_r), mask_i=_ptr(mask_i),
                          mask_stride_n=msn, mask_stride_f=msf, window=window.data_ptr(), mel_dev=_ptr(mel_dev),
                          out=out.data_ptr(), stats=stats.data_ptr(), tables=plan.tables(dev),
                          l1_target=_ptr(l1_target), l1_rows=_ptr(l1_rows), frame_lens=frame_lens.data_ptr(),
                          cuda_stream=torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.aas_lmfb_forward_ex(plan.handle, ctypes.byref(io)))
        ctx.plan, ctx.flags, ctx.eps, ctx.tmax, ctx.n_ch = plan, flags, float(eps), tmax, n_ch
        ctx.strides = (msn, msf)
        ctx.save_for_backward(wave, lengths, mask_r, mask_i, window, out, stats, mel_dev)
        ctx.mark_non_differentiable(frame_lens)
        if l1_rows is not None:
            ctx.mark_non_differentiable(l1_rows)
            return out, frame_lens, l1_rows
        return out, frame_lens

    @staticmethod
    def backward(ctx, grad_out, _grad_lens, _grad_rows=None):
        wave, lengths, mask_r, mask_i, window, out, stats, mel_dev = ctx.saved_tensors
        want_wave = ctx.needs_input_grad[0]
        if mask_r is None and not want_wave:
            return (None,) * 12
        lib = _lib.load()
        dev = wave.device
        n, plan, tmax, n_ch = wave.shape[0], ctx.plan, ctx.tmax, ctx.n_ch
        lmax = wave.shape[-1]
        grad_out = grad_out.contiguous()
        if grad_out.dtype != torch.float32:
            grad_out = grad_out.float()
        gr = torch.empty_strided(mask_r.shape, mask_r.stride(), dtype=torch.float32, device=dev) \
            if mask_r is not None else None
        gi = torch.empty_strided(mask_i.shape, mask_i.stride(), dtype=torch.float32, device=dev) \
            if mask_i is not None else None
        ws_bytes = lib.aas_lmfb_workspace_bytes(plan.handle, n, tmax, ctx.flags)
        ws = torch.empty((max(ws_bytes, 4) + 3) // 4, dtype=torch.float32, device=dev)
        msn, msf = ctx.strides
        gw = None
        if want_wave:
            # gradient into the samples too (a waveform-domain enhancer in front of this op).  The C ABI
            # has one set of strides for wave and grad_wave: zeros_like would return contiguous strides
            # for a wave that is a slice of a larger buffer
            gw = torch.empty_strided(wave.shape, wave.stride(), dtype=torch.float32, device=dev) \
                if _dense_strides(wave) else None
            if gw is None:
                wave = wave.contiguous()
                gw = torch.empty_like(wave)
            if HOP * tmax < lmax:                          # (the library zeroes the first min(Lmax, 160 tmax) samples of every
                gw.zero_()                                 #  row; only a caller-chosen smaller tmax leaves a tail)
        io = _lib.make_io(flags=ctx.flags, device=dev.index if dev.index is not None else -1, n=n, n_ch=n_ch, tmax=tmax,
                          eps=ctx.eps, wave=wave.data_ptr(), wave_stride=_row_stride(wave, n_ch),
                          wave_stride_ch=wave.stride(1) if wave.dim() == 3 else 0, wave_len=lmax,
                          lengths=lengths.data_ptr(), mask_r=_ptr(mask_r), mask_i=_ptr(mask_i),
                          mask_stride_n=msn, mask_stride_f=msf, window=window.data_ptr(), mel_dev=_ptr(mel_dev),
                          out=out.data_ptr(), stats=stats.data_ptr(), grad_out=grad_out.data_ptr(),
                          grad_mask_r=_ptr(gr), grad_mask_i=_ptr(gi), grad_wave=_ptr(gw), workspace=ws.data_ptr(),
                          tables=plan.tables(dev), cuda_stream=torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.aas_lmfb_backward_ex(plan.handle, ctypes.byref(io)))
        return gw, None, gr, gi, None, None, None, None, None, None, None, None


def _dense_strides(wave):
    """Rows (and channels) of the wave do not overlap, so a gradient buffer with the SAME strides can
    be allocated (a slice of a larger buffer qualifies, an expanded view does not)."""
    lmax = wave.shape[-1]
    if wave.dim() == 2:
        return wave.shape[0] == 1 or wave.stride(0) >= lmax
    n, c = wave.shape[0], wave.shape[1]
    if c > 1 and wave.stride(1) < lmax:
        return False
    return n == 1 or wave.stride(0) >= (c - 1) * wave.stride(1) + lmax


def _row_stride(wave, n_ch):
    """Utterance stride handed to the C ABI; with a single utterance it is never multiplied by
    anything but zero, so an even value keeps the 8-byte staging path available."""
    if wave.shape[0] > 1:
        return wave.stride(0)
    tot = n_ch * wave.shape[-1] if wave.dim() == 2 else (n_ch - 1) * wave.stride(1) + wave.shape[-1]
    return tot + (tot & 1)


class LMFBFrontEnd(torch.nn.Module):
    """``nn.Module`` face of :class:`LMFB`: holds ``mel_basis`` and ``window`` as buffers the
    way ``BRNNmultiCH`` holds ``mel_basis`` (model.py:167).

    ``forward(wave, lengths, mask_r=None, mask_i=None)`` -> ``(features (N, n_mels, Tmax),
    frame_lens)``.  With ``mask_mode='reim'`` it is the tail of ``BRNNmultiCH.forward``
    (model.py:186-198) fused with the STFT in front of it and the CMVN behind it.
    """

    def __init__(self, mel_basis=None, window=None, mask_mode="reim", cmvn_mode="per_bin",
                 eps=0.0, n_mels=40, window_sym=True):
        super().__init__()
        mel = slaney_mel_basis(n_mels=n_mels) if mel_basis is None else mel_basis
        mel = torch.as_tensor(np.asarray(mel.detach().cpu() if isinstance(mel, torch.Tensor) else mel),
                              dtype=torch.float32)
        win = hamming_window(N_FFT, window_sym) if window is None else window
        win = torch.as_tensor(np.asarray(win.detach().cpu() if isinstance(win, torch.Tensor) else win),
                              dtype=torch.float32)
        if win.numel() != N_FFT:
            raise ValueError("window must have 320 samples (AM_training/train.py:39-40)")
        self.register_buffer("mel_basis", mel)
        self.register_buffer("window", win)
        self.mask_mode, self.cmvn_mode, self.eps = mask_mode, cmvn_mode, float(eps)
        self._plan, self._plan_key = None, None
        self._tuning = (0, 0, False)

    def set_tuning(self, warps_fwd=0, warps_bwd=0, static_schedule=False):
        """Tests / benchmarks: warps per 32-frame tile of the forward / backward kernel (4..6, 0 = the
        measured default) and the static tile schedule.  Results do not depend on them."""
        self._tuning = (int(warps_fwd), int(warps_bwd), int(bool(static_schedule)))
        self._plan_key = None
        return self

    @property
    def plan(self):
        """The native plan of the CURRENT ``mel_basis`` buffer: rebuilt when the buffer is replaced
        (``.to()``) or written in place (``load_state_dict`` copies into it), so a checkpoint with a
        different basis can never run on a stale plan."""
        mel = self.mel_basis
        key = (id(mel), mel._version, tuple(mel.shape))
        if self._plan is None or key != self._plan_key:
            host = mel.detach().cpu()
            if self._plan is None or host.shape != self._plan.mel.shape or \
                    not np.array_equal(host.numpy(), self._plan.mel):
                self._plan = MelPlan(host)
            _lib.check(self._plan._lib.aas_lmfb_plan_set_tuning(self._plan.handle, *[int(v) for v in self._tuning]))
            self._plan_key = key
        return self._plan

    @property
    def audio_conf(self):
        """The front-end configuration, to be stored next to checkpoints (cf. the unused
        ``audio_conf`` slot of DeepSpeech.serialize, model.py:267, :432-433)."""
        return dict(sample_rate=SAMPLE_RATE, window_size=N_FFT / SAMPLE_RATE,
                    window_stride=HOP / SAMPLE_RATE, window="hamming",
                    n_mels=int(self.mel_basis.shape[0]), mask_mode=self.mask_mode,
                    cmvn_mode=self.cmvn_mode, eps=self.eps)

    def forward(self, wave, lengths, mask_r=None, mask_i=None, tmax=None, l1_target=None):
        """-> ``(features, frame_lens)``; with ``l1_target`` (the ``target`` of the ``L1Loss_mask`` that
        follows, e.g. the clean features in trainer_DCE.py) -> ``(features, frame_lens, l1_rows)``: the
        per-row sums ``sum_t |Z - target|`` formed by the CMVN kernel while Z is in registers; hand them to
        ``L1Loss_mask()(features, target, mask, rows=l1_rows)`` and Z is not read again for the loss."""
        return LMFB.apply(wave, lengths, mask_r, mask_i, self.plan, self.window,
                          self.mask_mode, self.cmvn_mode, self.eps, tmax, self.mel_basis, l1_target)

    @torch.no_grad()
    def stft(self, wave, lengths, tmax=None):
        """The enhancer's input: ``(N, 2*161, Tmax)`` with the 161 real rows first and the 161
        imaginary rows second -- the layout ``BRNNmultiCH.forward`` views as ``(N, 2, F, T)``
        (model.py:170, :186-188) -- and ``frame_lens``.  Same framing, window and FFT as
        :meth:`forward`; frames ``t >= T_i`` are exact zeros.  Not differentiable (it is data)."""
        if not wave.is_cuda:
            raise RuntimeError("LMFB is CUDA-only (sm_100a); there is no CPU fallback")
        dev = wave.device
        _check_f32_cuda("wave", wave, dev)
        if wave.dim() != 2 or wave.stride(1) != 1:
            raise ValueError("wave must be (N, Lmax) with unit sample stride")
        lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
        n = wave.shape[0]
        tmax = int(1 + wave.shape[1] // HOP if tmax is None else tmax)
        out = torch.empty((n, 2 * N_BINS, tmax), dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.aas_lmfb_stft(self.plan.handle, wave.data_ptr(), lengths.data_ptr(), n, wave.stride(0),
                                   self.window.data_ptr(), out.data_ptr(), out.stride(0), tmax, stream)
        _lib.check(rc)
        frame_lens = torch.clamp(1 + torch.div(lengths, HOP, rounding_mode="floor"), max=tmax)
        frame_lens = torch.where(lengths >= 1, frame_lens, torch.zeros_like(frame_lens)).to(torch.int32)
        return out, frame_lens

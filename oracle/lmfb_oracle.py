"""CPU oracle for the LMFB front-end hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or the CPU
baseline -- never as the thing shipped.  The product (``aas_enhancement_b200``) must not
import anything from ``oracle/``.

What is restated here, and from where (paths relative to ``/root/reference``):

* front-end parameters (16 kHz, 20 ms Hamming window, 10 ms stride, 40 mels, n_fft = 320,
  161 one-sided bins)                      ``AM_training/train.py:39-42``, ``:55-56``, ``:199``
* mask -> power -> mel -> log1p glue       ``Speech_enhancement_by_AAS/model.py:186-198``
* the mel basis as a caller-supplied ``(40, F)`` matrix applied as a k=1 conv
                                            ``Speech_enhancement_by_AAS/model.py:148``, ``:167``, ``:196``
* length-sorted zero-pad collate, byte mask (1 = padding), float32 ``input_percentages``
                                            ``Speech_enhancement_by_AAS/loader_functions.py:47-105``
* downstream length arithmetic (float32 multiply then truncation)
                                            ``Speech_enhancement_by_AAS/trainer_AAS.py:165-167``
* "log(1+S) + CMVN" / ``--normalize``      ``AM_training/train.py:59-61``

PARITY PINNING.  The mask->power->mel->log1p glue and the collate/length code are pinned
against the *live* reference (``tests/golden/make_golden.py`` imports ``model.py`` and
``loader_functions.py`` from ``/root/reference`` and the fixtures it wrote are committed).
The STFT front (framing, reflect padding, window symmetry, default mel values) and the
normalisation statistic live in un-vendored code (``data.data_loader.SpectrogramDataset``,
imported at ``AM_training/train.py:11`` but absent from the tree; derived from
SeanNaren/deepspeech.pytorch on top of librosa, ``AM_training/requirements.txt:5``, no
version pinned).  For those steps this oracle restates the published librosa /
deepspeech.pytorch algorithms; the reference holds no golden vectors for them, so at that
boundary the status is **parity unpinned**.

Everything is float64 unless stated; the DFT is ``numpy.fft.rfft``.
"""
from __future__ import annotations

import math

import numpy as np

SAMPLE_RATE = 16000          # AM_training/train.py:39
WINDOW_SIZE_S = 0.02         # AM_training/train.py:40
WINDOW_STRIDE_S = 0.01       # AM_training/train.py:41
N_FFT = int(SAMPLE_RATE * WINDOW_SIZE_S)       # 320
HOP = int(SAMPLE_RATE * WINDOW_STRIDE_S)       # 160
N_BINS = N_FFT // 2 + 1      # 161, AM_training/train.py:199
N_MELS = 40                  # AM_training/train.py:56, config.py:41


# --------------------------------------------------------------------------- parameters
def hamming_window(n: int = N_FFT, sym: bool = True) -> np.ndarray:
    """Hamming window (name pinned by AM_training/train.py:42).

    ``sym=True`` is scipy.signal.hamming as deepspeech.pytorch passes it to librosa;
    ``sym=False`` is the periodic (DFT-even) variant.
    """
    k = np.arange(n, dtype=np.float64)
    denom = (n - 1) if sym else n
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / denom)


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr: int = SAMPLE_RATE, n_fft: int = N_FFT, n_mels: int = N_MELS,
                   fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """Slaney-scale, area-normalised triangular filterbank, shape (n_mels, n_fft//2+1).

    The published ``librosa.filters.mel`` algorithm (htk=False, norm='slaney'), which is
    what the reference's missing SpectrogramDataset would have used.  The reference itself
    only pins that the basis is a caller-supplied (40, F) matrix (model.py:148, :167).
    """
    if fmax is None:
        fmax = sr / 2.0
    n_bins = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    weights *= enorm[:, None]
    return weights


def frame_count(n_samples, hop: int = HOP):
    """Frames of a centred STFT: T = 1 + L // hop (101/401/601/3001 for 1/4/6/30 s)."""
    return 1 + np.asarray(n_samples) // hop


# --------------------------------------------------------------------------- STFT front
def reflect_index(i: np.ndarray, length: int) -> np.ndarray:
    """Index map of numpy 'reflect' padding (edge sample not repeated), iterated so that
    it is defined for any offset; one bounce is all a >=161-sample signal ever needs."""
    if length == 1:
        return np.zeros_like(i)
    period = 2 * (length - 1)
    j = np.mod(i, period)
    return np.where(j >= length, period - j, j)


def stft_frames(wave: np.ndarray, length: int, window: np.ndarray,
                n_fft: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """Centred STFT of one utterance; returns complex (F, T) with T = 1 + length // hop.

    Each utterance is reflect-padded by n_fft//2 on both sides **using its own length**
    (librosa.stft(center=True, pad_mode='reflect')).
    """
    x = np.asarray(wave, dtype=np.float64)[:length]
    t = int(frame_count(length, hop))
    idx = (np.arange(t)[:, None] * hop + np.arange(n_fft)[None, :]) - n_fft // 2
    frames = x[reflect_index(idx, length)] * window[None, :]
    return np.fft.rfft(frames, axis=1).T        # (F, T)


# --------------------------------------------------------------------------- forward
def masked_power(re, im, mask_r, mask_i, mask_mode: str):
    """model.py:191-194 for 'reim'; 'power' and 'none' as named in BASELINE.json."""
    if mask_mode == "reim":
        return (re * mask_r) ** 2 + (im * mask_i) ** 2
    if mask_mode == "power":
        return mask_r * (re ** 2 + im ** 2)
    if mask_mode == "none":
        return re ** 2 + im ** 2
    raise ValueError(mask_mode)


def cmvn(y: np.ndarray, mode: str, eps: float = 0.0):
    """Per-utterance mean/variance normalisation of y (M, T) over its own T frames.

    'per_bin': mean/std per mel bin over time; 'global': scalar mean/std over the whole
    matrix (deepspeech.pytorch ``normalize``); both use the unbiased std (torch .std()).
    Returns (z, mean, rstd) with mean/rstd of shape (M,).
    """
    m, t = y.shape
    if mode == "none":
        return y.copy(), np.zeros(m), np.ones(m)
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode == "per_bin":
            mean = y.mean(axis=1)
            var = ((y - mean[:, None]) ** 2).sum(axis=1) / (t - 1)
        elif mode == "global":
            mean = np.full(m, y.mean())
            var = np.full(m, ((y - y.mean()) ** 2).sum() / (y.size - 1))
        else:
            raise ValueError(mode)
        rstd = 1.0 / (np.sqrt(var) + eps)
        z = (y - mean[:, None]) * rstd[:, None]
    return z, mean, rstd


def lmfb_forward(wave, lengths, mask_r=None, mask_i=None, mel=None, window=None,
                 mask_mode: str = "reim", cmvn_mode: str = "per_bin", eps: float = 0.0,
                 return_parts: bool = False):
    """LMFB-v0 forward on a zero-padded batch.

    wave (N, Lmax), lengths (N,), masks (N, F, Tmax) or None.  Returns Z (N, M, Tmax)
    float64 with frames t >= T_i exactly 0 (collate zero-fill happens after the
    per-utterance normalisation, loader_functions.py:56, :66) and frame_lens (N,) int32.

    Multi-channel (BRNNmultiCH with nCH > 1, model.py:160-167, :186-198): wave (N, nCH, Lmax),
    masks (N, nCH*F, Tmax) with channel c in rows [c*F, (c+1)*F); the mel basis repeats over the
    channels (``mel_basis.repeat(1, nCH)``, model.py:167), i.e. the masked powers are summed.
    """
    wave = np.asarray(wave, dtype=np.float64)
    lengths = np.asarray(lengths, dtype=np.int64)
    n = wave.shape[0]
    if mel is None:
        mel = mel_filterbank()
    if window is None:
        window = hamming_window()
    mel = np.asarray(mel, dtype=np.float64)
    window = np.asarray(window, dtype=np.float64)
    n_fft = window.shape[0]
    hop = n_fft // 2
    frame_lens = frame_count(lengths, hop).astype(np.int32)
    if mask_r is not None:
        tmax = mask_r.shape[2]
    else:
        tmax = int(frame_lens.max()) if n else 0
    z = np.zeros((n, mel.shape[0], tmax), dtype=np.float64)
    parts = []
    for i in range(n):
        t_i = int(frame_lens[i])
        chans = wave[i] if wave.ndim == 3 else wave[i][None]
        n_bins = n_fft // 2 + 1
        p = 0.0
        for c in range(chans.shape[0]):
            spec = stft_frames(chans[c], int(lengths[i]), window, n_fft, hop)
            re, im = spec.real, spec.imag
            rows = slice(c * n_bins, (c + 1) * n_bins)
            mr = None if mask_r is None else np.asarray(mask_r[i, rows, :t_i], dtype=np.float64)
            mi = None if mask_i is None else np.asarray(mask_i[i, rows, :t_i], dtype=np.float64)
            p = p + masked_power(re, im, mr, mi, mask_mode)
        e = mel @ p                                   # model.py:196 (k=1 conv == matmul; basis repeated over channels :167)
        y = np.log1p(e)                               # model.py:198
        zi, mean, rstd = cmvn(y, cmvn_mode, eps)
        z[i, :, :t_i] = zi
        if return_parts:
            parts.append(dict(re=re, im=im, p=p, e=e, y=y, mean=mean, rstd=rstd))
    if return_parts:
        return z, frame_lens, parts
    return z, frame_lens


# --------------------------------------------------------------------------- autograd twin
def lmfb_forward_torch(wave, lengths, mask_r=None, mask_i=None, mel=None, window=None,
                       mask_mode: str = "reim", cmvn_mode: str = "per_bin", eps: float = 0.0,
                       dtype=None):
    """Differentiable torch twin of :func:`lmfb_forward` (float64 by default).

    Uses an explicit DFT matrix (no torch.stft) so that gradients flow to masks and wave.
    """
    import torch

    dtype = dtype or torch.float64
    if mel is None:
        mel = mel_filterbank()
    if window is None:
        window = hamming_window()
    mel_t = torch.as_tensor(np.asarray(mel), dtype=dtype)
    win_t = torch.as_tensor(np.asarray(window), dtype=dtype)
    n_fft = win_t.shape[0]
    hop = n_fft // 2
    n_bins = n_fft // 2 + 1
    k = np.arange(n_bins)[:, None] * np.arange(n_fft)[None, :]
    ang = -2.0 * np.pi * k / n_fft
    cos_m = torch.as_tensor(np.cos(ang), dtype=dtype)
    sin_m = torch.as_tensor(np.sin(ang), dtype=dtype)
    lengths = [int(v) for v in lengths]
    n = len(lengths)
    frame_lens = [1 + v // hop for v in lengths]
    tmax = mask_r.shape[2] if mask_r is not None else max(frame_lens)
    rows = []
    for i in range(n):
        t_i = frame_lens[i]
        idx = (np.arange(t_i)[:, None] * hop + np.arange(n_fft)[None, :]) - n_fft // 2
        idx = torch.as_tensor(reflect_index(idx, lengths[i]))
        chans = wave[i] if wave.dim() == 3 else wave[i][None]
        p = 0.0
        for c in range(chans.shape[0]):
            frames = chans[c].to(dtype)[idx] * win_t[None, :]     # (T, n_fft)
            re = cos_m @ frames.T                                 # (F, T)
            im = sin_m @ frames.T
            fr = slice(c * n_bins, (c + 1) * n_bins)
            if mask_mode == "reim":
                p = p + (re * mask_r[i, fr, :t_i].to(dtype)) ** 2 + (im * mask_i[i, fr, :t_i].to(dtype)) ** 2
            elif mask_mode == "power":
                p = p + mask_r[i, fr, :t_i].to(dtype) * (re ** 2 + im ** 2)
            else:
                p = p + re ** 2 + im ** 2
        y = torch.log1p(mel_t @ p)
        if cmvn_mode == "per_bin":
            y = (y - y.mean(dim=1, keepdim=True)) / (y.std(dim=1, keepdim=True) + eps)
        elif cmvn_mode == "global":
            y = (y - y.mean()) / (y.std() + eps)
        rows.append(torch.nn.functional.pad(y, (0, tmax - t_i)))
    return torch.stack(rows), torch.tensor(frame_lens, dtype=torch.int32)


def lmfb_grads(wave, lengths, mask_r, mask_i, grad_out, mel=None, window=None,
               mask_mode: str = "reim", cmvn_mode: str = "per_bin", eps: float = 0.0,
               want_wave_grad: bool = False):
    """float64 gradients of sum(Z * grad_out) w.r.t. the masks (and optionally the wave)."""
    import torch

    w = torch.as_tensor(np.asarray(wave), dtype=torch.float64)
    if want_wave_grad:
        w.requires_grad_(True)
    mr = mi = None
    leaves = []
    if mask_mode in ("reim", "power"):
        mr = torch.as_tensor(np.asarray(mask_r), dtype=torch.float64).requires_grad_(True)
        leaves.append(mr)
    if mask_mode == "reim":
        mi = torch.as_tensor(np.asarray(mask_i), dtype=torch.float64).requires_grad_(True)
        leaves.append(mi)
    if want_wave_grad:
        leaves.append(w)
    z, _ = lmfb_forward_torch(w, lengths, mr, mi, mel, window, mask_mode, cmvn_mode, eps)
    g = torch.as_tensor(np.asarray(grad_out), dtype=torch.float64)
    (z * g).sum().backward()
    out = {"z": z.detach().numpy()}
    if mr is not None:
        out["grad_mask_r"] = mr.grad.numpy()
    if mi is not None:
        out["grad_mask_i"] = mi.grad.numpy()
    if want_wave_grad:
        out["grad_wave"] = w.grad.numpy()
    return out


# --------------------------------------------------------------------------- glue only
def glue_reim(stft_real, stft_imag, mask_real, mask_imag, mel_basis):
    """Exactly model.py:191-198 on (N, F, T) arrays: the one piece the reference pins."""
    enh_real = stft_real * mask_real
    enh_imag = stft_imag * mask_imag
    enh_power = enh_real ** 2 + enh_imag ** 2
    enh_mel = np.einsum("mf,nft->nmt", mel_basis, enh_power)
    return np.log1p(enh_mel)


# --------------------------------------------------------------------------- batch layout
def collate(batch):
    """loader_functions.py:47-73 restated for a list of (tensor (C, T_i), target list).

    Returns (inputs (N, C, Tmax) f32, targets i32, input_percentages f32, target_sizes i32,
    mask (N, 1, Tmax) u8 with 1 = padding), sorted by T descending (stable).
    """
    batch = sorted(batch, key=lambda s: s[0].shape[1], reverse=True)
    c = batch[0][0].shape[0]
    tmax = batch[0][0].shape[1]
    n = len(batch)
    inputs = np.zeros((n, c, tmax), dtype=np.float32)
    pct = np.zeros(n, dtype=np.float32)
    tsz = np.zeros(n, dtype=np.int32)
    mask = np.zeros((n, 1, tmax), dtype=np.uint8)
    targets = []
    for x, (feat, target) in enumerate(batch):
        t = feat.shape[1]
        inputs[x, :, :t] = feat
        pct[x] = t / float(tmax)                  # double divide, stored as float32 (:57)
        tsz[x] = len(target)
        targets.extend(target)
        if t < tmax:
            mask[x, :, t:] = 1
    return inputs, np.asarray(targets, dtype=np.int32), pct, tsz, mask


def collate_paired(batch):
    """loader_functions.py:75-105 for (noisy (C,T), txt list, clean (C,T)) triples."""
    batch = sorted(batch, key=lambda s: s[0].shape[1], reverse=True)
    c = batch[0][0].shape[0]
    tmax = batch[0][0].shape[1]
    n = len(batch)
    inputs = np.zeros((n, c, tmax), dtype=np.float32)
    outputs = np.zeros((n, c, tmax), dtype=np.float32)
    pct = np.zeros(n, dtype=np.float32)
    tsz = np.zeros(n, dtype=np.int32)
    mask = np.zeros((n, 1, tmax), dtype=np.uint8)
    targets = []
    for x, (feat, txt, clean) in enumerate(batch):
        t = feat.shape[1]
        inputs[x, :, :t] = feat
        outputs[x, :, :t] = clean                 # clean padded to the NOISY length (:97)
        if t < tmax:
            mask[x, :, t:] = 1
        pct[x] = t / float(tmax)
        tsz[x] = len(txt)
        targets.extend(txt)
    return inputs, outputs, mask, np.asarray(targets, dtype=np.int32), pct, tsz


def conv_out_frames(t: int, kernel: int = 11, stride: int = 2, n_down: int = 1) -> int:
    """Time length after the 2-conv DeepSpeech front (model.py:289-297): first conv has the
    given stride, the second stride 1 when nDownsample == 1."""
    t1 = (t - kernel) // stride + 1
    s2 = 1 if n_down == 1 else stride
    return (t1 - kernel) // s2 + 1


def ctc_sizes(input_percentages: np.ndarray, t_out: int) -> np.ndarray:
    """trainer_AAS.py:165-167: float32 multiply by int(T') then truncation to int32."""
    pct = np.asarray(input_percentages, dtype=np.float32)
    return (pct * np.float32(int(t_out))).astype(np.int32)


def l1loss_mask(input, target, mask, fix_masking: bool = False):
    """Speech_enhancement_by_AAS/model.py:19-31 restated: returns (loss, nElement).

    The reference's ``err.masked_fill(mask, 0)`` is not in-place (a no-op), so padded frames
    contribute; the divisor is the number of unmasked FRAMES ``numel(mask) - sum(mask)``.
    ``fix_masking=True`` applies the mask for real (opt-in, not the reference's behaviour)."""
    mask = np.asarray(mask)
    if mask[0][0][0] != 0:
        raise RuntimeError("nElement is undefined in the reference when mask[0][0][0] != 0")
    n_element = mask.size - int(mask.sum())
    err = np.abs(np.asarray(input, dtype=np.float64) - np.asarray(target, dtype=np.float64))
    if fix_masking:
        err = np.where(mask.astype(bool), 0.0, err)
    return err.sum() / n_element, n_element


def rel_err(a, b) -> float:
    """max|a-b| / max(|b|, rms(b)) elementwise (hard part 4)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    rms = math.sqrt(float(np.mean(b * b)))
    denom = np.maximum(np.abs(b), rms if rms > 0 else 1.0)
    return float(np.max(np.abs(a - b) / denom))

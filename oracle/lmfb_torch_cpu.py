"""The reference's CPU feature path restated with stock torch CPU ops.  TEST/BASELINE ONLY.

This is the ``cpu_baseline`` / ``bench.py --impl reference`` arm (kind "port"): the path a
user of the reference would run on the host -- ``torch.stft(320, hop 160, hamming, centre,
reflect)`` (standing in for the librosa STFT of the missing SpectrogramDataset,
``AM_training/train.py:11``, ``:190-199``), then the *literal* ops of
``Speech_enhancement_by_AAS/model.py:191-198`` (``torch.mul``, ``torch.pow`` + ``torch.pow``,
``F.conv1d(power, mel_basis[40,161,1])``, ``torch.log1p``), per-utterance mean/std
normalisation (``AM_training/train.py:59-61``) and autograd ``backward`` into the masks.

``torch.stft`` is allowed here and nowhere on the product path.  fp32, as the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

N_FFT = 320
HOP = 160


def forward_batched(wave, mask_r, mask_i, mel_basis, window, cmvn_mode="per_bin",
                    mask_mode="reim"):
    """Equal-length batch (best case for the CPU): wave (N, L) -> (N, 40, T)."""
    spec = torch.stft(wave, N_FFT, hop_length=HOP, win_length=N_FFT, window=window,
                      center=True, pad_mode="reflect", return_complex=True)   # (N, F, T)
    stft_real, stft_imag = spec.real, spec.imag
    if mask_mode == "reim":
        enh_real = torch.mul(stft_real, mask_r)                 # model.py:191
        enh_imag = torch.mul(stft_imag, mask_i)                 # model.py:192
        enh_power = torch.pow(enh_real, 2) + torch.pow(enh_imag, 2)   # model.py:194
    elif mask_mode == "power":
        enh_power = mask_r * (torch.pow(stft_real, 2) + torch.pow(stft_imag, 2))
    else:
        enh_power = torch.pow(stft_real, 2) + torch.pow(stft_imag, 2)
    enh_mel = F.conv1d(enh_power, mel_basis.unsqueeze(-1))      # model.py:196
    out = torch.log1p(enh_mel)                                  # model.py:198
    if cmvn_mode == "per_bin":
        out = (out - out.mean(dim=2, keepdim=True)) / out.std(dim=2, keepdim=True)
    elif cmvn_mode == "global":
        out = (out - out.mean(dim=(1, 2), keepdim=True)) / out.std(dim=(1, 2), keepdim=True)
    return out


def fwd_bwd_batched(wave, mask_r, mask_i, grad_out, mel_basis, window, cmvn_mode="per_bin",
                    mask_mode="reim"):
    """One training-style step: forward, then backward into both masks."""
    mr = mask_r.detach().requires_grad_(True) if mask_r is not None else None
    mi = mask_i.detach().requires_grad_(True) if mask_i is not None else None
    out = forward_batched(wave, mr, mi, mel_basis, window, cmvn_mode, mask_mode)
    leaves = [t for t in (mr, mi) if t is not None]
    if leaves:
        out.backward(grad_out)
    return out.detach(), (mr.grad if mr is not None else None), (mi.grad if mi is not None else None)


def fwd_bwd_per_utterance(wave, lengths, mask_r, mask_i, grad_out, mel_basis, window,
                          cmvn_mode="per_bin"):
    """The reference's actual usage pattern: one STFT per utterance, as a DataLoader worker
    would run it (AM_training/train.py:255-268, num_workers=1 :35)."""
    outs = []
    for i in range(wave.shape[0]):
        t_i = 1 + int(lengths[i]) // HOP
        o, _, _ = fwd_bwd_batched(wave[i:i + 1, :int(lengths[i])], mask_r[i:i + 1, :, :t_i],
                                  mask_i[i:i + 1, :, :t_i], grad_out[i:i + 1, :, :t_i],
                                  mel_basis, window, cmvn_mode)
        outs.append(o)
    return outs

"""Wave-level batch layout: the host-side mirror of the reference's collate functions.

The reference batches *precomputed features*; here the same tuples are produced from raw
waveforms so that the front-end can run on the GPU inside the training step:

* ``collate_wave``        <-> ``_collate_fn``        (loader_functions.py:47-73)
* ``collate_wave_paired`` <-> ``_collate_fn_paired`` (loader_functions.py:75-105)
* ``ctc_sizes``           <-> ``input_percentages.mul_(int(T)).int()`` (trainer_AAS.py:165-167)
* ``get_variable_nograd`` <-> ``_get_variable_nograd`` (utils.py:155-160)

Same tuple order, dtypes and sort order as the reference; the only difference is that
``inputs`` holds zero-padded waves ``(N, Lmax)`` instead of features ``(N, 40, Tmax)`` and a
trailing ``lengths`` (samples, int32) is appended.  Feature lengths are ``T_i = 1 + L_i // 160``.
"""
from __future__ import annotations

import torch

HOP = 160


def frame_count(n_samples: int) -> int:
    """Frames of the centred 320/160 STFT (AM_training/train.py:39-41): 1 + L // 160."""
    return 1 + int(n_samples) // HOP


def _as_wave(x) -> torch.Tensor:
    w = torch.as_tensor(x, dtype=torch.float32)
    if w.dim() != 1:
        raise ValueError("each wave must be 1-D (mono)")
    return w


def collate_wave(batch, pin_memory: bool = False):
    """batch: list of ``(wave (L_i,), target list[int])``.

    Returns ``(inputs, targets, input_percentages, target_sizes, mask, lengths)`` -- the first
    five exactly as loader_functions.py:73: sorted by feature length descending (stable),
    ``input_percentages[x] = T_x / float(Tmax)`` stored as float32, ``target_sizes`` int32,
    flat int32 ``targets``, ``mask (N, 1, Tmax)`` uint8 with 1 = padding.
    """
    batch = sorted(batch, key=lambda s: frame_count(len(s[0])), reverse=True)
    n = len(batch)
    lmax = max(len(s[0]) for s in batch)
    tmax = frame_count(len(batch[0][0]))
    inputs = torch.zeros(n, lmax, pin_memory=pin_memory)
    input_percentages = torch.FloatTensor(n)
    target_sizes = torch.IntTensor(n)
    lengths = torch.IntTensor(n)
    mask = torch.zeros(n, 1, tmax, dtype=torch.uint8)
    targets = []
    for x, sample in enumerate(batch):
        wave, target = _as_wave(sample[0]), sample[1]
        li = wave.numel()
        seq_length = frame_count(li)
        inputs[x, :li].copy_(wave)
        input_percentages[x] = seq_length / float(tmax)
        target_sizes[x] = len(target)
        lengths[x] = li
        targets.extend(target)
        if seq_length < tmax:
            mask[x, :, seq_length:].fill_(1)
    targets = torch.IntTensor(targets)
    return inputs, targets, input_percentages, target_sizes, mask, lengths


def collate_wave_paired(batch, pin_memory: bool = False):
    """batch: list of ``(noisy wave, txt list[int], clean wave)``.

    Returns ``(inputs, outputs, mask, targets, input_percentages, target_sizes, lengths)`` in
    the order of loader_functions.py:105; ``outputs`` (clean) is laid out to the NOISY length
    (loader_functions.py:85, :97).
    """
    batch = sorted(batch, key=lambda s: frame_count(len(s[0])), reverse=True)
    n = len(batch)
    lmax = max(len(s[0]) for s in batch)
    tmax = frame_count(len(batch[0][0]))
    inputs = torch.zeros(n, lmax, pin_memory=pin_memory)
    outputs = torch.zeros(n, lmax, pin_memory=pin_memory)
    mask = torch.zeros(n, 1, tmax, dtype=torch.uint8)
    input_percentages = torch.FloatTensor(n)
    target_sizes = torch.IntTensor(n)
    lengths = torch.IntTensor(n)
    targets = []
    for x, sample in enumerate(batch):
        wave, txt, clean = _as_wave(sample[0]), sample[1], _as_wave(sample[2])
        li = wave.numel()
        seq_length = frame_count(li)
        inputs[x, :li].copy_(wave)
        lc = min(li, clean.numel())
        outputs[x, :lc].copy_(clean[:lc])
        if seq_length < tmax:
            mask[x, :, seq_length:].fill_(1)
        input_percentages[x] = seq_length / float(tmax)
        target_sizes[x] = len(txt)
        lengths[x] = li
        targets.extend(txt)
    targets = torch.IntTensor(targets)
    return inputs, outputs, mask, targets, input_percentages, target_sizes, lengths


def ctc_sizes(input_percentages: torch.Tensor, t_out: int) -> torch.Tensor:
    """trainer_AAS.py:165-167: float32 multiply by ``int(T')`` then truncation.  This is NOT
    ``floor(T * T' / Tmax)`` in exact arithmetic (e.g. 135/435 * 203 -> 62, not 63), so the
    float32 path is reproduced literally (out of place, unlike the reference's ``mul_``)."""
    return input_percentages.to(torch.float32).mul(int(t_out)).int()


def shard_utterances(n: int, world_size: int, rank: int):
    """Indices of a length-sorted batch owned by ``rank``: round-robin, which balances total
    frames (30 utterances over 8 ranks -> 4,4,4,4,4,4,3,3).  No data-path collective."""
    return list(range(rank, n, world_size))


def get_variable_nograd(inputs: torch.Tensor, cuda: bool = True, non_blocking: bool = True):
    """utils.py:155-160 without the removed ``Variable`` wrapper: the host->device boundary."""
    out = inputs.cuda(non_blocking=non_blocking) if cuda else inputs
    return out.requires_grad_(False) if out.is_floating_point() else out

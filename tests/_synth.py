"""Seeded synthetic inputs shared by the tests and tests/golden/make_golden.py.

numpy's legacy RandomState stream is frozen across numpy versions, so fixtures made here
can be regenerated bit-for-bit.  Seed 123 is the reference's default seed (config.py:62).
"""
from __future__ import annotations

import numpy as np

HOP = 160
N_BINS = 161
N_MELS = 40


def make_batch(n, max_len, seed=123, ragged=False, tonal=False, min_frac=0.5, mask_lo=0.0,
               lengths=None):
    """wave (n, max_len) f32 zero-padded, lengths i32, mask_r/mask_i (n,161,Tmax) f32,
    grad_out (n,40,Tmax) f32.  Lengths sorted descending like the reference collate."""
    rs = np.random.RandomState(seed)
    if lengths is None:
        if ragged:
            lengths = (rs.uniform(min_frac, 1.0, size=n) * max_len).astype(np.int64)
            lengths[0] = max_len
            lengths = np.sort(lengths)[::-1].copy()
        else:
            lengths = np.full(n, max_len, dtype=np.int64)
    lengths = np.asarray(lengths, dtype=np.int64)
    max_len = int(max(max_len, lengths.max()))
    wave = np.zeros((n, max_len), dtype=np.float32)
    for i in range(n):
        li = int(lengths[i])
        if tonal:
            t = np.arange(li) / 16000.0
            x = np.zeros(li)
            for _ in range(5):
                f = rs.uniform(50.0, 7900.0)
                a = rs.uniform(0.02, 0.2)
                ph = rs.uniform(0, 2 * np.pi)
                x += a * np.sin(2 * np.pi * f * t + ph)
            x += 0.01 * rs.randn(li)
        else:
            x = 0.1 * rs.randn(li)
        wave[i, :li] = np.clip(x, -1.0, 1.0).astype(np.float32)
    tmax = 1 + max_len // HOP
    mask_r = rs.uniform(mask_lo, 1.0, size=(n, N_BINS, tmax)).astype(np.float32)
    mask_i = rs.uniform(mask_lo, 1.0, size=(n, N_BINS, tmax)).astype(np.float32)
    grad_out = rs.randn(n, N_MELS, tmax).astype(np.float32)
    return dict(wave=wave, lengths=lengths.astype(np.int32), mask_r=mask_r, mask_i=mask_i,
                grad_out=grad_out, tmax=tmax)

"""aas_enhancement_b200 -- B200-native LMFB front-end for lifelongeek/AAS_enhancement.

Only the data-parallel hot path lives here: waveform -> STFT(320/160, Hamming) -> mask ->
40-band mel -> log1p -> CMVN, forward and backward, as hand-written sm_100a kernels behind a
C ABI (``include/aas_lmfb.h``), plus the host-side mirror of the reference's batch layout.
"""
from .lmfb import (LMFB, LMFBFrontEnd, MelPlan, hamming_window, slaney_mel_basis,
                   N_FFT, HOP, N_BINS)
from .losses import L1Loss_mask, CTCLoss
from .collate import (collate_wave, collate_wave_paired, ctc_sizes, frame_count,
                      shard_utterances, get_variable_nograd)
from .loader import (WaveDataset, WaveSampler, WaveLoader, WaveLoader_paired, WaveDataLoader,
                     load_wave, save_wave, to_device)

__all__ = ["LMFB", "LMFBFrontEnd", "MelPlan", "L1Loss_mask", "CTCLoss", "hamming_window", "slaney_mel_basis",
           "collate_wave", "collate_wave_paired", "ctc_sizes", "frame_count",
           "shard_utterances", "get_variable_nograd", "WaveDataset", "WaveSampler", "WaveLoader",
           "WaveLoader_paired", "WaveDataLoader", "load_wave", "save_wave", "to_device",
           "N_FFT", "HOP", "N_BINS"]

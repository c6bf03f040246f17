// Host-side analysis of a mel basis into the banded-2 form the kernels consume.
// Plain C++ (no CUDA) so that the CPU emulation harness can share it.
#pragma once
#include "lmfb_core.cuh"

namespace aas_lmfb {

// mel: (n_mels, kBins) row-major.  Returns 0 on success, -1 if some bin feeds a filter outside
// the two live ones (basis not banded / not frequency-ordered).
inline int build_mel_band(const float* mel, int n_mels, MelBand* out) {
    int hi[kMaxMels];                 // last bin with a non-zero weight, per filter (-1: empty)
    for (int m = 0; m < n_mels; ++m) {
        hi[m] = -1;
        for (int f = 0; f < kBins; ++f) if (mel[m * kBins + f] != 0.0f) hi[m] = f;
    }
    out->n_mels = (uint8_t)n_mels;
    int ml = 0;
    for (int f = 0; f < kBins; ++f) {
        while (ml < n_mels && hi[ml] < f) ++ml;              // lowest filter not yet finished
        for (int m = 0; m < n_mels; ++m)
            if (mel[m * kBins + f] != 0.0f && (m < ml || m > ml + 1)) return -1;
        out->ml[f] = (uint8_t)ml;
        out->wl[f] = ml < n_mels ? 0.25f * mel[ml * kBins + f] : 0.0f;
        out->wh[f] = ml + 1 < n_mels ? 0.25f * mel[(ml + 1) * kBins + f] : 0.0f;
    }
    return 0;
}

}  // namespace aas_lmfb

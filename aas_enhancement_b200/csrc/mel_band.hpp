// Host-side analysis of a mel basis into the banded-2 form the kernels consume.
// Plain C++ (no CUDA) so that the CPU emulation harness can share it.
#pragma once
#include "lmfb_core.cuh"

namespace aas_lmfb {

// mel: (n_mels, kBins) row-major.  Returns 0 on success, -1 if some bin feeds a filter outside
// the two live ones (basis not banded / not frequency-ordered).
inline uint32_t kBinOffHost(int f) {
    if (f == 0) return 0u;
    if (f == kBins - 1) return 1u;
    return (uint32_t)(((f % 5) * 32 + (f & 31)) * kPitch * 2);
}

inline int build_mel_band(const float* mel, int n_mels, MelBand* out, int* ml_out = nullptr) {
    int hi[kMaxMels];                 // last bin with a non-zero weight, per filter (-1: empty)
    for (int m = 0; m < n_mels; ++m) {
        hi[m] = -1;
        for (int f = 0; f < kBins; ++f) if (mel[m * kBins + f] != 0.0f) hi[m] = f;
    }
    out->n_mels = n_mels;
    for (int m = 0; m < kMaxMels; ++m) out->fend[m] = 0;
    int ml = 0;
    for (int f = 0; f < kBins; ++f) {
        while (ml < n_mels && hi[ml] < f) ++ml;              // lowest filter not yet finished
        for (int m = 0; m < n_mels; ++m)
            if (mel[m * kBins + f] != 0.0f && (m < ml || m > ml + 1)) return -1;
        if (ml_out) ml_out[f] = ml;
        BinEnt& e = out->ent[f];
        e.wl = ml < n_mels ? 0.25f * mel[ml * kBins + f] : 0.0f;
        e.wh = ml + 1 < n_mels ? 0.25f * mel[(ml + 1) * kBins + f] : 0.0f;
        e.off = kBinOffHost(f);
        e.sel = f == 0 ? 1u : (f == kBins - 1 ? 2u : 0u);
        for (int m = ml; m < n_mels; ++m) out->fend[m] = (uint8_t)(f + 1);
    }
    // now fend[m] = one past the last bin whose lower filter is <= m (empty ranges repeat the value)
    return 0;
}

}  // namespace aas_lmfb

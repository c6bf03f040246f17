// Host-side analysis of a mel basis into the banded-2 form the kernels consume.
// Plain C++ (no CUDA) so that the CPU emulation harness can share it.
#pragma once
#include "lmfb_core.cuh"

namespace aas_lmfb {

inline uint32_t kBinOffHost(int f) {
    if (f == 0) return 0u;
    if (f == kBins - 1) return 1u;
    return (uint32_t)(((f % 5) * 32 + (f & 31)) * kPitch * 2);
}

// mel: (n_mels, kBins) row-major.  `warps` = warps per tile of the forward kernel (phase-3 split).
// Returns 0 on success, -1 if some bin feeds a filter outside the two live ones (basis not
// banded / not frequency-ordered).  ml_out (optional, kBins ints) receives the lower filter of
// every bin (n_mels for bins above the last filter).
inline void split_filters(MelBand* out, int warps);

// Per-launch patch: row offsets depend on the caller's strides.
//   fwd: ent[f].moff = f * msf * 4 bytes.   bwd: additionally ent[f].off = dlo[f] * sem * 4 bytes.
// msf <= 2^22 and sem <= 2^22 keep both below 2^32.
inline void patch_strides(MelBand* band, unsigned msf, const uint8_t* dlo /*nullable*/, unsigned sem) {
    for (int f = 0; f < kBins; ++f) {
        band->ent[f].moff = (uint32_t)f * msf * 4u;
        if (dlo) band->ent[f].off = (uint32_t)dlo[f] * sem * 4u;
    }
}

// Backward table from the forward one: weights re-expressed on the always-valid row pair
// (dlo, dlo+1) of dE.  Needs n_mels >= 2.
inline void make_bwd_band(const MelBand& fwd, const int* ml, MelBand* bwd, uint8_t* dlo) {
    *bwd = fwd;
    const int n_mels = fwd.n_mels;
    for (int f = 0; f < kBins; ++f) {
        BinEnt& e = bwd->ent[f];
        if (ml[f] <= n_mels - 2)      { dlo[f] = (uint8_t)ml[f]; }
        else if (ml[f] == n_mels - 1) { dlo[f] = (uint8_t)(n_mels - 2); e.wh = fwd.ent[f].wl; e.wl = 0.0f; }
        else                          { dlo[f] = 0; e.wl = e.wh = 0.0f; }
    }
}

inline int build_mel_band(const float* mel, int n_mels, int warps, MelBand* out, int* ml_out = nullptr) {
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
        e.moff = 0;
        for (int m = ml; m < n_mels; ++m) out->fend[m] = (uint8_t)(f + 1);
    }
    // now fend[m] = one past the last bin whose lower filter is <= m (empty ranges repeat the value)

    split_filters(out, warps);
    return 0;
}

// split the filters between the warps of the forward kernel, balancing bins*6 + 40 per filter
inline void split_filters(MelBand* out, int warps) {
    const int n_mels = out->n_mels;
    if (warps < 1) warps = 1;
    if (warps > kMaxW) warps = kMaxW;
    int cost[kMaxMels], total = 0;
    for (int m = 0; m < n_mels; ++m) {
        const int lo = m > 0 ? out->fend[m - 1] : 0;
        cost[m] = 6 * (out->fend[m] - lo) + 40;
        total += cost[m];
    }
    for (int w = 0; w <= kMaxW; ++w) out->mbeg[w] = (uint8_t)n_mels;
    out->mbeg[0] = 0;
    int m = 0, acc = 0;
    for (int w = 1; w < warps; ++w) {
        const int target = (int)((long long)total * w / warps);
        while (m < n_mels && acc + cost[m] / 2 < target) acc += cost[m++];
        out->mbeg[w] = (uint8_t)m;
    }
}

}  // namespace aas_lmfb

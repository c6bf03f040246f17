// Host-side analysis of a mel basis into the banded-2 tables the kernels consume.
// Plain C++ (no CUDA) so that the CPU emulation harness can share it.
#pragma once
#include <string.h>
#include "lmfb_core.cuh"

namespace aas_lmfb {

// mel: (n_mels, kBins) row-major.  Returns 0 on success, -1 if some bin feeds a filter outside the
// two live ones (basis not banded / not frequency-ordered).  ml_out (kBins ints) receives the
// lower filter of every bin (n_mels for bins above the last filter).
inline int build_fwd_tab(const float* mel, int n_mels, FwdTab* out, int* ml_out) {
    int hi[kMaxMels];                 // last bin with a non-zero weight, per filter (-1: empty)
    for (int m = 0; m < n_mels; ++m) {
        hi[m] = -1;
        for (int f = 0; f < kBins; ++f) if (mel[m * kBins + f] != 0.0f) hi[m] = f;
    }
    memset(out, 0, sizeof(*out));
    out->n_mels = n_mels;
    int ml = 0;
    for (int f = 0; f < kBins; ++f) {
        const int before = ml;
        while (ml < n_mels && hi[ml] < f) ++ml;              // lowest filter not yet finished
        for (int m = 0; m < n_mels; ++m)
            if (mel[m * kBins + f] != 0.0f && (m < ml || m > ml + 1)) return -1;
        ml_out[f] = ml;
        out->w[f].x = ml < n_mels ? 0.25f * mel[ml * kBins + f] : 0.0f;
        out->w[f].y = ml + 1 < n_mels ? 0.25f * mel[(ml + 1) * kBins + f] : 0.0f;
        const int adv = f > 0 ? ml - before : 0;             // filters completed before bin f
        out->adv[f] = (uint8_t)adv;
        if (adv != 0) out->hmask[f >> 3] |= (uint8_t)(1u << (f & 7));
        if (adv > 1) out->multi = 1;
    }
    return 0;
}

// forward phase 3 with W warps: warp w walks the 8-bin groups [p3_g0(W, w), p3_g1(W, w)), the last
// warp also bin 160, and produces partial sums for filters lo[w] .. hi[w] = ml(first bin) ..
// ml(last bin) + 1
inline void set_warp_ranges(FwdTab* tab, const int* ml, int warps) {
    if (warps < 1) warps = 1;
    if (warps > kMaxW) warps = kMaxW;
    for (int w = 0; w < kMaxW; ++w) {
        int b0 = p3_g0(warps, w) * 8, b1 = p3_g1(warps, w) * 8;              // bins [b0, b1)
        if (w >= warps) b0 = b1 = 0;
        if (w == warps - 1) { if (b0 >= b1) b0 = kBins - 1; b1 = kBins; }
        tab->lo[w] = (uint8_t)(b0 < b1 ? ml[b0] : 255);
        tab->hi[w] = (uint8_t)(b0 < b1 ? ml[b1 - 1] + 1 : 0);
    }
}

// Backward table from the forward one: per pass-2 step k2 and output k1, the weights of bins
// f = (96 k1 + 65 k2) mod 160 and fp = 160 - f re-expressed on the always-valid row pair
// (d, d + 1) of dE.  Needs n_mels >= 2.
inline void build_bwd_tab(const FwdTab& fwd, const int* ml, BwdTab* bwd) {
    memset(bwd, 0, sizeof(*bwd));
    const int n_mels = fwd.n_mels;
    bwd->n_mels = n_mels;
    for (int k2 = 0; k2 < 17; ++k2)
        for (int k1 = 0; k1 < 5; ++k1)
            for (int side = 0; side < 2; ++side) {
                const int f0 = bin_of(k2, k1), f = side ? kBins - 1 - f0 : f0;
                float wl = fwd.w[f].x, wh = fwd.w[f].y;
                int d;
                if (ml[f] <= n_mels - 2)      { d = ml[f]; }
                else if (ml[f] == n_mels - 1) { d = n_mels - 2; wh = wl; wl = 0.0f; }
                else                          { d = 0; wl = wh = 0.0f; }
                bwd->w[k2][k1][2 * side] = wl;
                bwd->w[k2][k1][2 * side + 1] = wh;
                bwd->d[k2][k1][side] = (uint32_t)d;
            }
}

}  // namespace aas_lmfb

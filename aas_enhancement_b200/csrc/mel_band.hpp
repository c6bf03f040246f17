// Host-side analysis of a mel basis into the tables the kernels consume.
// Plain C++ (no CUDA) so that the CPU emulation harness can share it.
//
// The reference applies ANY (M, F) matrix with a k=1 conv1d (model.py:167, :196).  Here:
//   forward  : "walkable" = banded-2 with a non-decreasing lower filter (every triangular filterbank):
//              the bins are walked once with two running sums.  Otherwise every filter is gathered
//              over its row [lo, hi] with the weights read from the caller's device copy of the matrix.
//   backward : "banded" = every bin feeds at most two ADJACENT filters (again every triangular
//              filterbank, in any frequency order of the bins): dP[f] = wl dE[d] + wh dE[d+1] straight
//              from the table.  Otherwise the caller of the kernel first forms dP = 1/4 B^T dE
//              (162 rows) and the kernel runs with the identity table below.
#pragma once
#include <string.h>
#include "lmfb_core.cuh"

namespace aas_lmfb {

// mel: (n_mels, kBins) row-major.  Fills the forward table: the row descriptors always, and the
// walk tables when the basis is banded-2 with a non-decreasing lower filter (out->walkable).
inline void build_fwd_tab(const float* mel, int n_mels, FwdTab* out, int* ml_out) {
    int hi[kMaxMels], lo[kMaxMels];   // last / first bin with a non-zero weight, per filter (hi = -1: empty)
    for (int m = 0; m < n_mels; ++m) {
        hi[m] = -1; lo[m] = kBins;
        for (int f = 0; f < kBins; ++f) if (mel[m * kBins + f] != 0.0f) { if (f < lo[m]) lo[m] = f; hi[m] = f; }
    }
    memset(out, 0, sizeof(*out));
    out->n_mels = n_mels;
    for (int m = 0; m < n_mels; ++m)
        out->row[m] = hi[m] < 0 ? 0u : ((uint32_t)lo[m] | ((uint32_t)(hi[m] - lo[m] + 1) << 8));
    out->walkable = 1;
    int ml = 0;
    for (int f = 0; f < kBins; ++f) {
        const int before = ml;
        while (ml < n_mels && hi[ml] < f) ++ml;              // lowest filter not yet finished
        for (int m = 0; m < n_mels; ++m)
            if (mel[m * kBins + f] != 0.0f && (m < ml || m > ml + 1)) out->walkable = 0;
        ml_out[f] = ml;
        out->w[f].x = ml < n_mels ? 0.25f * mel[ml * kBins + f] : 0.0f;
        out->w[f].y = ml + 1 < n_mels ? 0.25f * mel[(ml + 1) * kBins + f] : 0.0f;
        const int adv = f > 0 ? ml - before : 0;             // filters completed before bin f
        out->adv[f] = (uint8_t)adv;
        if (adv != 0) out->hmask[f >> 3] |= (uint8_t)(1u << (f & 7));
        if (adv > 1) out->multi = 1;
    }
}

// forward phase 3 with W warps: the filters are dealt to the warps as W contiguous runs, warp w OWNS
// filters lo[w] .. hi[w] and walks every bin that feeds one of them (the bins whose lower filter is
// lo - 1 .. hi; ml is non-decreasing, so they are one range [b0, b1)).  The partition minimises the
// largest per-warp cost, cost = kP3CostBin per walked bin + kP3CostFilter per owned filter (log1p +
// store, done four filters at a time): the low filters of a mel basis have one or two bins each, the
// high ones ten and more.
#ifndef LMFB_P3_COST_BIN
#  define LMFB_P3_COST_BIN 8
#endif
#ifndef LMFB_P3_COST_FILTER
#  define LMFB_P3_COST_FILTER 40
#endif
constexpr int kP3CostBin = LMFB_P3_COST_BIN, kP3CostFilter = LMFB_P3_COST_FILTER;

inline void set_warp_ranges(FwdTab* tab, const int* ml, int warps) {
    if (warps < 1) warps = 1;
    if (warps > kMaxW) warps = kMaxW;
    const int M = tab->n_mels;
    int below[kMaxMels + 2];                                  // below[m] = bins whose lower filter is < m
    for (int m = 0; m <= M + 1; ++m) {
        below[m] = 0;
        for (int f = 0; f < kBins; ++f) if (ml[f] < m) ++below[m];
    }
    auto cost = [&](int a, int b) {                           // filters a .. b (inclusive), a <= b
        return kP3CostBin * (below[b + 1] - below[a > 0 ? a - 1 : 0]) + kP3CostFilter * ((b - a + 4) / 4 * 4);   // (log1p in groups of four)
    };
    // best[k][j]: filters [0, j) in k runs (runs may be empty), smallest possible largest cost
    static thread_local int best[kMaxW + 1][kMaxMels + 1], cut[kMaxW + 1][kMaxMels + 1];
    for (int j = 0; j <= M; ++j) { best[0][j] = j == 0 ? 0 : 0x3fffffff; cut[0][j] = 0; }
    for (int k = 1; k <= warps; ++k)
        for (int j = 0; j <= M; ++j) {
            best[k][j] = best[k - 1][j]; cut[k][j] = j;       // run k empty
            for (int i = 0; i < j; ++i) {
                const int c = cost(i, j - 1), v = best[k - 1][i] > c ? best[k - 1][i] : c;
                if (v < best[k][j]) { best[k][j] = v; cut[k][j] = i; }
            }
        }
    int end = M;
    for (int w = kMaxW - 1; w >= 0; --w) {
        int a = 1, b = 0;                                     // none
        if (w < warps) { a = cut[w + 1][end]; b = end - 1; end = a; }
        tab->lo[w] = (uint8_t)a; tab->hi[w] = (uint8_t)b;
        int b0 = 0, b1 = 0;
        if (a <= b) { b0 = below[a > 0 ? a - 1 : 0]; b1 = below[b + 1]; }
        tab->b0[w] = (uint8_t)b0; tab->b1[w] = (uint8_t)b1;
        tab->m0[w] = (uint8_t)(b0 < b1 ? ml[b0] : 1);
        tab->m1[w] = (uint8_t)(b0 < b1 ? ml[b1 - 1] + 1 : 0);
    }
}

// Backward table (per pass-2 step k2 and output k1: bins f = (96 k1 + 65 k2) mod 160 and fp = 160 - f),
// weights re-expressed on the always-valid row pair (d, d + 1) of dE and carrying the 1/4 that undoes
// X' = 2X.  Returns 0, or -1 when some bin feeds filters that are not one or two adjacent rows
// (then use build_bwd_tab_identity with a precomputed dP).  Needs n_mels >= 2.
inline int build_bwd_tab(const float* mel, int n_mels, BwdTab* bwd) {
    memset(bwd, 0, sizeof(*bwd));
    bwd->n_mels = n_mels;
    for (int k2 = 0; k2 < 17; ++k2)
        for (int k1 = 0; k1 < 5; ++k1)
            for (int side = 0; side < 2; ++side) {
                const int f0 = bin_of(k2, k1), f = side ? kBins - 1 - f0 : f0;
                int first = -1, last = -1;
                for (int m = 0; m < n_mels; ++m)
                    if (mel[m * kBins + f] != 0.0f) { if (first < 0) first = m; last = m; }
                float wl = 0.0f, wh = 0.0f;
                int d = 0;
                if (first >= 0) {
                    if (last - first > 1) return -1;
                    if (first <= n_mels - 2) {
                        d = first;
                        wl = 0.25f * mel[first * kBins + f];
                        wh = 0.25f * mel[(first + 1) * kBins + f];
                    } else {                                   // only the last filter: rows (M-2, M-1)
                        d = n_mels - 2;
                        wh = 0.25f * mel[first * kBins + f];
                    }
                }
                bwd->w[k2][k1][2 * side] = wl;
                bwd->w[k2][k1][2 * side + 1] = wh;
                bwd->d[k2][k1][side] = (uint32_t)d;
            }
    return 0;
}

// The kernel reads dP itself: a tensor of kBins + 1 rows whose row f is dP[f] (row 161 is never
// weighted), see dp_generic in lmfb_kernels.cu.
constexpr int kDpRows = kBins + 1;
inline void build_bwd_tab_identity(BwdTab* bwd) {
    memset(bwd, 0, sizeof(*bwd));
    bwd->n_mels = kDpRows;
    for (int k2 = 0; k2 < 17; ++k2)
        for (int k1 = 0; k1 < 5; ++k1)
            for (int side = 0; side < 2; ++side) {
                const int f0 = bin_of(k2, k1), f = side ? kBins - 1 - f0 : f0;
                bwd->w[k2][k1][2 * side] = 1.0f;
                bwd->d[k2][k1][side] = (uint32_t)f;
            }
}

}  // namespace aas_lmfb

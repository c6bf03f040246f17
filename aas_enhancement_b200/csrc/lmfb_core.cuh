// Per-thread building blocks of the fused STFT -> mask -> mel (-> log1p) kernels.
//
// Execution model ("lane = frame"): one warp owns a tile of 32 consecutive frames of one
// utterance and lane t does the whole 320-point real FFT of frame t0+t by itself, using a
// private column of shared memory as scratch.  Consequences:
//   * every twiddle is a literal (the FFT code is identical for all lanes),
//   * every global access of a (N, F, T)/(N, M, T) tensor has the lanes along T, i.e. is a
//     128-byte coalesced row segment -- no transposition is ever needed,
//   * no block-level synchronisation (only __syncwarp after staging).
//
// Real FFT of 320 samples = complex FFT of the 160 packed samples z[j] = x[2j] + i x[2j+1],
// computed with the Good-Thomas prime-factor map 160 = 5 x 32 (no inter-stage twiddles):
//   pass 1: five 32-point FFTs in registers (generated codelet), in place in the scratch;
//   pass 2: 5-point DFTs for the index pair (k2, 32-k2) followed by the real-split
//           butterfly, in place; bin f ends up in slot (f mod 5)*32 + (f mod 32), bins 0
//           and 160 (both real) share slot 0.
// The spectrum kept in the scratch is X' = 2 X (the 1/2 of the split is folded into the mel
// weights as 1/4, exact in binary floating point).
//
// The same code compiles as plain C++ for the CPU emulation harness under tests/emu/ (test
// infrastructure only).
//
// Reference semantics being reproduced: Speech_enhancement_by_AAS/model.py:186-198
// (mask -> power -> mel -> log1p), AM_training/train.py:39-42,:199 (320/160/hamming, 161 bins).
#pragma once
#include <stdint.h>
#include "fft_codelets.cuh"

#ifdef __CUDACC__
#  define LMFB_LDG(p) __ldg(p)
#else
#  define LMFB_LDG(p) (*(p))
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace aas_lmfb {

constexpr int kNfft  = 320;
constexpr int kHop   = 160;
constexpr int kBins  = 161;
constexpr int kTile  = 32;             // frames per warp tile
constexpr int kPitch = 33;             // float2 per scratch slot (32 lanes + 1 pad: conflict-free staging)
constexpr int kSlots = 160;
constexpr int kMaxMels = 128;
constexpr int kScratchBytes = kSlots * kPitch * 8;     // 42,240 B per warp

enum MaskMode { kMaskNone = 0, kMaskReim = 1, kMaskPower = 2 };

// Banded-2 description of the mel basis, passed by value as a kernel parameter (constant
// bank).  Bin f feeds filters ml[f] (weight wl[f]) and ml[f]+1 (weight wh[f]); ml is
// non-decreasing.  Weights already carry the 1/4 that undoes X' = 2X.
struct MelBand {
    float   wl[kBins];
    float   wh[kBins];
    uint8_t ml[kBins];
    uint8_t n_mels;
    uint8_t pad_[2];
};

LMFB_HD int slot_of_packed(int j) {            // j in [0,160): packed-sample index -> PFA input slot
    const int n1 = (3 * (j % 5)) % 5;
    const int n2 = (13 * (j & 31)) & 31;
    return n1 * 32 + n2;
}

LMFB_HD int slot_of_bin(int f) {               // f in [1,159]
    return (f % 5) * 32 + (f & 31);
}

// 'reflect' padding index (numpy semantics, any offset, length >= 1)
LMFB_HD int reflect_index(int i, int len) {
    if (len <= 1) return 0;
    const int period = 2 * (len - 1);
    int j = i % period;
    if (j < 0) j += period;
    return j >= len ? period - j : j;
}

// ---------------------------------------------------------------------------------------
// Staging: copy the (32+1)*160 samples a tile needs into the 32 private frame columns,
// windowed and permuted into PFA input order.  lanes run along the packed-sample index.
//   wave_row : first sample of this utterance;  len : its length (>=1)
//   t0       : first frame of the tile
//   S        : warp scratch base (float2 [kSlots][kPitch])
//   vec_ok   : wave_row is 8-byte aligned
// ---------------------------------------------------------------------------------------
LMFB_HD void stage_tile(int lane, const float* __restrict__ wave_row, int len, int t0,
                        const float* __restrict__ window, float2* __restrict__ S, bool vec_ok) {
    int   slot_a[3], slot_b[3];
    float wa0[3], wa1[3], wb0[3], wb1[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int c = lane + 32 * q;                       // packed index inside a hop-row, < 80 valid
        const int cc = c < 80 ? c : 0;
        slot_a[q] = slot_of_packed(cc) * kPitch;
        slot_b[q] = slot_of_packed(cc + 80) * kPitch;
        wa0[q] = LMFB_LDG(window + 2 * cc);
        wa1[q] = LMFB_LDG(window + 2 * cc + 1);
        wb0[q] = LMFB_LDG(window + 2 * cc + kHop);
        wb1[q] = LMFB_LDG(window + 2 * cc + kHop + 1);
    }
#pragma unroll 3
    for (int r = 0; r <= kTile; ++r) {
        const int row = t0 + r - 1;                        // hop-row index in the unpadded signal
        const long long base = (long long)row * kHop;
        const bool interior = vec_ok && row >= 0 && (base + kHop) <= (long long)len;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int c = lane + 32 * q;
            if (c < 80) {
                float2 v;
                if (interior) {
                    v = LMFB_LDG(reinterpret_cast<const float2*>(wave_row + base) + c);
                } else {
                    const int i0 = (int)base + 2 * c;
                    v.x = LMFB_LDG(wave_row + reflect_index(i0, len));
                    v.y = LMFB_LDG(wave_row + reflect_index(i0 + 1, len));
                }
                if (r < kTile)  S[slot_a[q] + r]     = make_float2(v.x * wa0[q], v.y * wa1[q]);
                if (r >= 1)     S[slot_b[q] + r - 1] = make_float2(v.x * wb0[q], v.y * wb1[q]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// pass 1: five in-register 32-point FFTs over the thread's private column (col = S + lane)
// ---------------------------------------------------------------------------------------
LMFB_HD void fft_pass1(float2* __restrict__ col) {
#pragma unroll 1
    for (int n1 = 0; n1 < 5; ++n1) {
        float2* p = col + n1 * 32 * kPitch;
        float xr[32], xi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float2 v = p[i * kPitch]; xr[i] = v.x; xi[i] = v.y; }
        fft32(xr, xi);
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i * kPitch] = make_float2(xr[i], xi[i]);
    }
}

LMFB_HD void dft5(const float (&ar)[5], const float (&ai)[5], float (&br)[5], float (&bi)[5]) {
    const float c1 = 3.090169944e-01f, c2 = -8.090169944e-01f;
    const float s1 = 9.510565163e-01f, s2 = 5.877852523e-01f;
    const float t1r = ar[1] + ar[4], t1i = ai[1] + ai[4];
    const float t2r = ar[2] + ar[3], t2i = ai[2] + ai[3];
    const float t3r = ar[1] - ar[4], t3i = ai[1] - ai[4];
    const float t4r = ar[2] - ar[3], t4i = ai[2] - ai[3];
    br[0] = ar[0] + t1r + t2r;
    bi[0] = ai[0] + t1i + t2i;
    const float m1r = fmaf(c2, t2r, fmaf(c1, t1r, ar[0])), m1i = fmaf(c2, t2i, fmaf(c1, t1i, ai[0]));
    const float m2r = fmaf(c1, t2r, fmaf(c2, t1r, ar[0])), m2i = fmaf(c1, t2i, fmaf(c2, t1i, ai[0]));
    const float n1r = fmaf(s2, t4r, s1 * t3r), n1i = fmaf(s2, t4i, s1 * t3i);
    const float n2r = fmaf(-s1, t4r, s2 * t3r), n2i = fmaf(-s1, t4i, s2 * t3i);
    br[1] = m1r + n1i; bi[1] = m1i - n1r;
    br[4] = m1r - n1i; bi[4] = m1i + n1r;
    br[2] = m2r + n2i; bi[2] = m2i - n2r;
    br[3] = m2r - n2i; bi[3] = m2i + n2r;
}

// real-split butterfly: A = Z[f], Bz = Z[160-f]; returns X'[f] and X'[160-f] (both scaled by 2)
LMFB_HD void split_pair(float Ar, float Ai, float Bzr, float Bzi, float sn, float cs,
                        float2& xf, float2& xp) {
    const float sr = Ar + Bzr, si = Ai - Bzi;        // S  = A + conj(Bz)
    const float dr = Ar - Bzr, di = Ai + Bzi;        // A - conj(Bz)
    const float Dr = fmaf(sn, dr, -cs * di);         // D' = (sin + i cos)(dr + i di)
    const float Di = fmaf(sn, di, cs * dr);
    xf = make_float2(sr - Dr, si - Di);
    xp = make_float2(sr + Dr, -(si + Di));
}

// ---------------------------------------------------------------------------------------
// pass 2: radix-5 across the five sub-transforms + real split, in place.
// ---------------------------------------------------------------------------------------
LMFB_HD void fft_pass2(float2* __restrict__ col) {
    float ar[5], ai[5], Ar[5], Ai[5], br[5], bi[5], Br[5], Bi[5];
    // ---- k2 = 0 (self-paired; holds bins 0, 160, and pairs (96,64), (32,128))
    {
#pragma unroll
        for (int n = 0; n < 5; ++n) { const float2 v = col[(n * 32) * kPitch]; ar[n] = v.x; ai[n] = v.y; }
        dft5(ar, ai, Ar, Ai);
        col[0] = make_float2(2.0f * (Ar[0] + Ai[0]), 2.0f * (Ar[0] - Ai[0]));   // X'[0], X'[160]
#pragma unroll
        for (int k1 = 1; k1 <= 2; ++k1) {
            float2 xf, xp;
            split_pair(Ar[k1], Ai[k1], Ar[5 - k1], Ai[5 - k1], kSplitSin[0][k1], kSplitCos[0][k1], xf, xp);
            col[(k1 * 32) * kPitch] = xf;
            col[((5 - k1) * 32) * kPitch] = xp;
        }
    }
    // ---- k2 = 16 (self-paired; bin 80 and pairs (16,144), (112,48))
    {
#pragma unroll
        for (int n = 0; n < 5; ++n) { const float2 v = col[(n * 32 + 16) * kPitch]; ar[n] = v.x; ai[n] = v.y; }
        dft5(ar, ai, Ar, Ai);
        col[16 * kPitch] = make_float2(2.0f * Ar[0], -2.0f * Ai[0]);             // X'[80] = 2 conj Z[80]
#pragma unroll
        for (int k1 = 1; k1 <= 2; ++k1) {
            float2 xf, xp;
            split_pair(Ar[k1], Ai[k1], Ar[5 - k1], Ai[5 - k1], kSplitSin[16][k1], kSplitCos[16][k1], xf, xp);
            col[(k1 * 32 + 16) * kPitch] = xf;
            col[((5 - k1) * 32 + 16) * kPitch] = xp;
        }
    }
    // ---- k2 = 1..15 paired with 32-k2
#pragma unroll 1
    for (int k2 = 1; k2 < 16; ++k2) {
        const int kb = 32 - k2;
#pragma unroll
        for (int n = 0; n < 5; ++n) {
            const float2 v = col[(n * 32 + k2) * kPitch]; ar[n] = v.x; ai[n] = v.y;
            const float2 w = col[(n * 32 + kb) * kPitch]; br[n] = w.x; bi[n] = w.y;
        }
        dft5(ar, ai, Ar, Ai);
        dft5(br, bi, Br, Bi);
#pragma unroll
        for (int k1 = 0; k1 < 5; ++k1) {
            const int kp = (5 - k1) % 5;
            float2 xf, xp;
            split_pair(Ar[k1], Ai[k1], Br[kp], Bi[kp], kSplitSin[k2][k1], kSplitCos[k2][k1], xf, xp);
            col[(k1 * 32 + k2) * kPitch] = xf;
            col[(kp * 32 + kb) * kPitch] = xp;
        }
    }
}

// (re', im') of bin f from the finished scratch column
LMFB_HD float2 load_bin(const float2* __restrict__ col, int f) {
    if (f == 0)          { const float2 v = col[0]; return make_float2(v.x, 0.0f); }
    if (f == kBins - 1)  { const float2 v = col[0]; return make_float2(v.y, 0.0f); }
    return col[slot_of_bin(f) * kPitch];
}

template <int MASK>
LMFB_HD float masked_power(float2 x, float mr, float mi) {
    if (MASK == kMaskReim) { const float a = x.x * mr, b = x.y * mi; return fmaf(a, a, b * b); }
    const float p = fmaf(x.x, x.x, x.y * x.y);
    return MASK == kMaskPower ? mr * p : p;
}

}  // namespace aas_lmfb

namespace aas_lmfb {

// ---------------------------------------------------------------------------------------
// phase 3 (forward): walk the bins in ascending order, multiply by the mask(s), accumulate the
// two live mel filters, and emit log1p(E[m]) whenever a filter is complete.
//   mr/mi   : mask_r/mask_i + n*stride_n + t      (row f is at + f*sf)
//   out     : out + n*stride_n + t                (row m is at + m*som)
//   inrow   : t < Tmax (memory exists);  valid : t < T_i (frame exists)
// ---------------------------------------------------------------------------------------
template <int MASK>
LMFB_HD void phase3_fwd(const float2* __restrict__ col, const MelBand& mb,
                        const float* __restrict__ mr, const float* __restrict__ mi, long long sf,
                        float* __restrict__ out, long long som, bool inrow, bool valid) {
    int m = 0;
    const int n_mels = mb.n_mels;
    float acc0 = 0.0f, acc1 = 0.0f;
#define LMFB_EMIT()                                                        \
    do {                                                                   \
        if (m < n_mels) {                                                  \
            const float y_ = valid ? log1pf(acc0) : 0.0f;                  \
            if (inrow) out[(long long)m * som] = y_;                       \
        }                                                                  \
        acc0 = acc1; acc1 = 0.0f; ++m;                                     \
    } while (0)
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int fb = 40 * b;
#pragma unroll
        for (int g = 0; g < 5; ++g) {
            float vr[8], vi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int f = fb + g * 8 + i;
                vr[i] = (MASK != kMaskNone && inrow) ? LMFB_LDG(mr + (long long)f * sf) : 0.0f;
                vi[i] = (MASK == kMaskReim && inrow) ? LMFB_LDG(mi + (long long)f * sf) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int fi = g * 8 + i;                  // compile-time
                const int f = fb + fi;
                float2 x = col[((fi % 5) * 32 + ((8 * b + fi) & 31)) * kPitch];
                if (fi == 0 && b == 0) x = make_float2(x.x, 0.0f);       // bin 0 is real (slot 0 .x)
                const float p = masked_power<MASK>(x, vr[i], vi[i]);
                const int ml = mb.ml[f];
                while (m < ml) LMFB_EMIT();
                acc0 = fmaf(mb.wl[f], p, acc0);
                acc1 = fmaf(mb.wh[f], p, acc1);
            }
        }
    }
    {   // bin 160 (real, slot 0 .y)
        const int f = kBins - 1;
        const float vr = (MASK != kMaskNone && inrow) ? LMFB_LDG(mr + (long long)f * sf) : 0.0f;
        const float vi = 0.0f;
        const float2 x = make_float2(col[0].y, 0.0f);
        const float p = masked_power<MASK>(x, vr, vi);
        const int ml = mb.ml[f];
        while (m < ml) LMFB_EMIT();
        acc0 = fmaf(mb.wl[f], p, acc0);
        acc1 = fmaf(mb.wh[f], p, acc1);
    }
    while (m < n_mels) LMFB_EMIT();
#undef LMFB_EMIT
}

// ---------------------------------------------------------------------------------------
// phase 3 (backward): dP[f] = wl*dE[ml] + wh*dE[ml+1];  'reim': dMr = 2 Mr Re^2 dP,
// dMi = 2 Mi Im^2 dP;  'power': dM = (Re^2 + Im^2) dP.
//   dE : dE + n*stride_n + t (row m at + m*sem), zero for frames t >= T_i
// ---------------------------------------------------------------------------------------
template <int MASK>
LMFB_HD void phase3_bwd(const float2* __restrict__ col, const MelBand& mb,
                        const float* __restrict__ mr, const float* __restrict__ mi, long long sf,
                        const float* __restrict__ dE, long long sem,
                        float* __restrict__ gr, float* __restrict__ gi, long long gsf, bool inrow) {
    int m = 0;
    const int n_mels = mb.n_mels;
    float d0 = (inrow && 0 < n_mels) ? LMFB_LDG(dE) : 0.0f;
    float d1 = (inrow && 1 < n_mels) ? LMFB_LDG(dE + sem) : 0.0f;
    float d2 = (inrow && 2 < n_mels) ? LMFB_LDG(dE + 2 * sem) : 0.0f;
#define LMFB_ADV()                                                                         \
    do {                                                                                   \
        d0 = d1; d1 = d2; ++m;                                                             \
        d2 = (inrow && m + 2 < n_mels) ? LMFB_LDG(dE + (long long)(m + 2) * sem) : 0.0f;   \
    } while (0)
#define LMFB_GRAD(x, vr_, vi_, f_)                                                         \
    do {                                                                                   \
        const int ml_ = mb.ml[f_];                                                         \
        while (m < ml_) LMFB_ADV();                                                        \
        const float dp_ = fmaf(mb.wh[f_], d1, mb.wl[f_] * d0);                             \
        if (MASK == kMaskReim) {                                                           \
            const float a_ = (x).x * (x).x * (vr_), b_ = (x).y * (x).y * (vi_);            \
            if (inrow) { gr[(long long)(f_) * gsf] = 2.0f * a_ * dp_;                      \
                         gi[(long long)(f_) * gsf] = 2.0f * b_ * dp_; }                    \
        } else {                                                                           \
            const float pw_ = fmaf((x).x, (x).x, (x).y * (x).y);                           \
            if (inrow) gr[(long long)(f_) * gsf] = pw_ * dp_;                              \
        }                                                                                  \
    } while (0)
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int fb = 40 * b;
#pragma unroll
        for (int g = 0; g < 5; ++g) {
            float vr[8], vi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int f = fb + g * 8 + i;
                vr[i] = (MASK == kMaskReim && inrow) ? LMFB_LDG(mr + (long long)f * sf) : 0.0f;
                vi[i] = (MASK == kMaskReim && inrow) ? LMFB_LDG(mi + (long long)f * sf) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int fi = g * 8 + i;
                const int f = fb + fi;
                float2 x = col[((fi % 5) * 32 + ((8 * b + fi) & 31)) * kPitch];
                if (fi == 0 && b == 0) x = make_float2(x.x, 0.0f);
                LMFB_GRAD(x, vr[i], vi[i], f);
            }
        }
    }
    {
        const int f = kBins - 1;
        const float vr = (MASK == kMaskReim && inrow) ? LMFB_LDG(mr + (long long)f * sf) : 0.0f;
        const float2 x = make_float2(col[0].y, 0.0f);
        LMFB_GRAD(x, vr, 0.0f, f);
    }
#undef LMFB_GRAD
#undef LMFB_ADV
}

}  // namespace aas_lmfb

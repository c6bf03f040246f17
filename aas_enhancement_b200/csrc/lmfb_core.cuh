// Per-thread building blocks of the fused STFT -> mask -> mel (-> log1p) kernels.
//
// Execution model ("lane = frame"): one warp owns a tile of 32 consecutive frames of one
// utterance and lane t does the whole 320-point real FFT of frame t0+t by itself, using a
// private column of shared memory as scratch.  Consequences:
//   * every twiddle is a literal (the FFT code is identical for all lanes),
//   * every global access of a (N, F, T)/(N, M, T) tensor has the lanes along T, i.e. is a
//     128-byte coalesced row segment -- no transposition is ever needed,
//   * no block-level synchronisation (only __syncwarp around staging).
//
// Real FFT of 320 samples = complex FFT of the 160 packed samples z[j] = x[2j] + i x[2j+1],
// computed with the Good-Thomas prime-factor map 160 = 5 x 32 (no inter-stage twiddles):
//   stage  : windowed samples -> private columns, permuted into PFA input order
//   pass 1 : five 32-point FFTs in registers (generated codelet), in place in the scratch
//   pass 2 : for each index pair (k2, 32-k2): two 5-point DFTs + the real-split butterfly give
//            ten bins; the mask(s) for those ten bins (prefetched one step ahead into
//            registers, lanes along T) are applied at once and the masked power (forward) or
//            the mask-gradient factors (backward) replace the spectrum in place.  Bin f lives
//            in slot (f mod 5)*32 + (f mod 32); bins 0 and 160 (both real) share slot 0.
//   phase 3: walk the bins in ascending order out of the scratch: banded mel accumulation +
//            log1p (forward), or dP * factors -> mask gradients (backward).
// The spectrum is kept as X' = 2 X (the 1/2 of the split is folded into the mel weights as
// 1/4, exact in binary floating point).
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
#  define LMFB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#  define LMFB_LDG(p) (*(p))
#  define LMFB_PREFETCH_L2(p) ((void)(p))
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
// bank).  Bin f feeds filters ml(f) (weight wl) and ml(f)+1 (weight wh) with ml non-decreasing,
// so the bins whose lower filter is m form the contiguous range [fend[m-1], fend[m]).
// Weights already carry the 1/4 that undoes X' = 2X.
struct BinEnt {
    float    wl, wh;
    uint32_t off;      // float offset of the bin's payload inside a scratch column
    uint32_t sel;      // 0: regular bin; 1: bin 0 (.x of slot 0, real); 2: bin 160 (.y of slot 0, real)
};
struct MelBand {
    BinEnt  ent[kBins];
    uint8_t fend[kMaxMels];
    int     n_mels;
};

LMFB_HD int slot_of_packed(int j) {            // j in [0,160): packed-sample index -> PFA input slot
    const int n1 = (3 * (j % 5)) % 5;
    const int n2 = (13 * (j & 31)) & 31;
    return n1 * 32 + n2;
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
// windowed and permuted into PFA input order.  Lanes run along the packed-sample index, so
// global reads are coalesced 256-byte runs and the scratch writes are conflict-free.
// Rows are loaded eight at a time (24 independent 8-byte loads in flight per lane).
// ---------------------------------------------------------------------------------------
struct StageLane {                      // per-lane constants of the staging map
    int   slot_a[3], slot_b[3];
    float wa0[3], wa1[3], wb0[3], wb1[3];
};

LMFB_HD void stage_lane_init(int lane, const float* __restrict__ window, StageLane& sl) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int c = lane + 32 * q;                       // packed index inside a hop-row, < 80 valid
        const int cc = c < 80 ? c : 0;
        sl.slot_a[q] = slot_of_packed(cc) * kPitch;
        sl.slot_b[q] = slot_of_packed(cc + 80) * kPitch;
        sl.wa0[q] = LMFB_LDG(window + 2 * cc);
        sl.wa1[q] = LMFB_LDG(window + 2 * cc + 1);
        sl.wb0[q] = LMFB_LDG(window + 2 * cc + kHop);
        sl.wb1[q] = LMFB_LDG(window + 2 * cc + kHop + 1);
    }
}

// rows [row_lo, row_hi) of the unpadded signal lie fully inside [0, len) and can be read as float2
LMFB_HD bool rows_interior(int row_lo, int row_hi, int len, bool vec_ok) {
    return vec_ok && row_lo >= 0 && (long long)row_hi * kHop <= (long long)len;
}

LMFB_HD void stage_store(const StageLane& sl, float2* __restrict__ S, int r, int q, float2 v) {
    if (r < kTile)  S[sl.slot_a[q] + r]     = make_float2(v.x * sl.wa0[q], v.y * sl.wa1[q]);
    if (r >= 1)     S[sl.slot_b[q] + r - 1] = make_float2(v.x * sl.wb0[q], v.y * sl.wb1[q]);
}

// edge rows (reflect padding at either end of the utterance, or an unaligned wave): one
// element at a time, kept out of line and rolled -- only boundary tiles ever come here
LMFB_HD void stage_rows_slow(int lane, const StageLane& sl, const float* __restrict__ wave_row, int len,
                             int t0, float2* __restrict__ S, int r_lo, int r_hi) {
#pragma unroll 1
    for (int r = r_lo; r < r_hi; ++r) {
        const int base = (t0 + r - 1) * kHop;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
            const int c = lane + 32 * q;
            if (c >= 80) continue;
            float2 v;
            v.x = LMFB_LDG(wave_row + reflect_index(base + 2 * c, len));
            v.y = LMFB_LDG(wave_row + reflect_index(base + 2 * c + 1, len));
            const int qq = q;                      // slot/window tables are tiny: index dynamically
            const int sa = qq == 0 ? sl.slot_a[0] : (qq == 1 ? sl.slot_a[1] : sl.slot_a[2]);
            const int sb = qq == 0 ? sl.slot_b[0] : (qq == 1 ? sl.slot_b[1] : sl.slot_b[2]);
            const float wa0 = qq == 0 ? sl.wa0[0] : (qq == 1 ? sl.wa0[1] : sl.wa0[2]);
            const float wa1 = qq == 0 ? sl.wa1[0] : (qq == 1 ? sl.wa1[1] : sl.wa1[2]);
            const float wb0 = qq == 0 ? sl.wb0[0] : (qq == 1 ? sl.wb0[1] : sl.wb0[2]);
            const float wb1 = qq == 0 ? sl.wb1[0] : (qq == 1 ? sl.wb1[1] : sl.wb1[2]);
            if (r < kTile)  S[sa + r]     = make_float2(v.x * wa0, v.y * wa1);
            if (r >= 1)     S[sb + r - 1] = make_float2(v.x * wb0, v.y * wb1);
        }
    }
}

LMFB_HD void stage_tile(int lane, const StageLane& sl, const float* __restrict__ wave_row, int len,
                        int t0, float2* __restrict__ S, bool vec_ok) {
    constexpr int kRowsPerBatch = 8;               // rows 0..31 in 4 batches, row 32 on its own
#pragma unroll 1
    for (int r0 = 0; r0 < kTile; r0 += kRowsPerBatch) {
        if (!rows_interior(t0 + r0 - 1, t0 + r0 - 1 + kRowsPerBatch, len, vec_ok)) {
            stage_rows_slow(lane, sl, wave_row, len, t0, S, r0, r0 + kRowsPerBatch);
            continue;
        }
        const float2* src = reinterpret_cast<const float2*>(wave_row + (long long)(t0 + r0 - 1) * kHop) + lane;
        float2 v[kRowsPerBatch][3];
#pragma unroll
        for (int i = 0; i < kRowsPerBatch; ++i) {
            v[i][0] = LMFB_LDG(src + i * 80);
            v[i][1] = LMFB_LDG(src + i * 80 + 32);
            v[i][2] = lane < 16 ? LMFB_LDG(src + i * 80 + 64) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int i = 0; i < kRowsPerBatch; ++i) {
            stage_store(sl, S, r0 + i, 0, v[i][0]);
            stage_store(sl, S, r0 + i, 1, v[i][1]);
            if (lane < 16) stage_store(sl, S, r0 + i, 2, v[i][2]);
        }
    }
    if (!rows_interior(t0 + kTile - 1, t0 + kTile, len, vec_ok)) {
        stage_rows_slow(lane, sl, wave_row, len, t0, S, kTile, kTile + 1);
    } else {
        const float2* src = reinterpret_cast<const float2*>(wave_row + (long long)(t0 + kTile - 1) * kHop) + lane;
        const float2 v0 = LMFB_LDG(src), v1 = LMFB_LDG(src + 32);
        const float2 v2 = lane < 16 ? LMFB_LDG(src + 64) : make_float2(0.0f, 0.0f);
        stage_store(sl, S, kTile, 0, v0);
        stage_store(sl, S, kTile, 1, v1);
        if (lane < 16) stage_store(sl, S, kTile, 2, v2);
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

// real-split butterfly: A = Z[f], Bz = Z[160-f]; returns X'[f] and X'[160-f] (both scaled by 2).
// Also correct for the self-paired bins: f = 0 gives (X'[0], X'[160]) and f = 80 gives X'[80] twice.
LMFB_HD void split_pair(float Ar, float Ai, float Bzr, float Bzi, float sn, float cs,
                        float2& xf, float2& xp) {
    const float sr = Ar + Bzr, si = Ai - Bzi;        // S  = A + conj(Bz)
    const float dr = Ar - Bzr, di = Ai + Bzi;        // A - conj(Bz)
    const float Dr = fmaf(sn, dr, -cs * di);         // D' = (sin + i cos)(dr + i di)
    const float Di = fmaf(sn, di, cs * dr);
    xf = make_float2(sr - Dr, si - Di);
    xp = make_float2(sr + Dr, -(si + Di));
}

// what pass 2 leaves in the scratch for one bin, given the spectrum value and the mask(s)
//   forward : .x = masked power P'            (.y unused)
//   backward: (.x, .y) = (dP -> dMr factor, dP -> dMi factor) = (2 Mr Re'^2, 2 Mi Im'^2)
//             'power' mode: .x = Re'^2 + Im'^2
template <int MASK, bool BWD>
LMFB_HD float2 bin_payload(float2 x, float mr, float mi) {
    if (!BWD) {
        if (MASK == kMaskReim) { const float a = x.x * mr, b = x.y * mi; return make_float2(fmaf(a, a, b * b), 0.0f); }
        const float p = fmaf(x.x, x.x, x.y * x.y);
        return make_float2(MASK == kMaskPower ? mr * p : p, 0.0f);
    }
    if (MASK == kMaskReim) return make_float2(2.0f * mr * x.x * x.x, 2.0f * mi * x.y * x.y);
    return make_float2(fmaf(x.x, x.x, x.y * x.y), 0.0f);
}

// which mask tensors a kernel variant reads
#define LMFB_NEEDS_MASK_R(MASK, BWD) ((BWD) ? (MASK) == kMaskReim : (MASK) != kMaskNone)
#define LMFB_NEEDS_MASK_I(MASK, BWD) ((MASK) == kMaskReim)

// mask values of the ten bins of pass-2 step k2: index k1 -> bin f, index 5+k1 -> bin 160-f.
// mr/mi point at a column that is always readable (out-of-row lanes are clamped by the caller),
// so the loads carry no predicate; sf (floats per mask row) fits 32 bits.
template <int MASK, bool BWD>
LMFB_HD void load_step_masks(int k2, const float* __restrict__ mr, const float* __restrict__ mi,
                             unsigned sf, float (&vr)[10], float (&vi)[10]) {
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const unsigned f = kBinOf[k2][k1];
        const unsigned of = f * sf, op = (kBins - 1 - f) * sf;
        vr[k1] = vr[5 + k1] = vi[k1] = vi[5 + k1] = 0.0f;
        if (LMFB_NEEDS_MASK_R(MASK, BWD)) { vr[k1] = LMFB_LDG(mr + of); vr[5 + k1] = LMFB_LDG(mr + op); }
        if (LMFB_NEEDS_MASK_I(MASK, BWD)) { vi[k1] = LMFB_LDG(mi + of); vi[5 + k1] = LMFB_LDG(mi + op); }
    }
}

// ---------------------------------------------------------------------------------------
// pass 2, one step: columns k2 and kb = (32-k2) mod 32 of the five sub-transforms -> two 5-point
// DFTs -> real split -> ten bins -> mask -> payload, in place.  Branch-free and identical for
// all k2: the self-paired columns (k2 = 0, 16, where kb == k2) simply compute each of their
// bins twice; only bins 0/160 (k2 = 0, k1 = 0), which share slot 0, need a select.
// ---------------------------------------------------------------------------------------
template <int MASK, bool BWD>
LMFB_HD void pass2_step(float2* __restrict__ col, int k2, const float (&vr)[10], const float (&vi)[10]) {
    const int kb = (32 - k2) & 31;
    float2* ca = col + k2 * kPitch;
    float2* cb = col + kb * kPitch;
    float ar[5], ai[5], br[5], bi[5], Ar[5], Ai[5], Br[5], Bi[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) {
        const float2 v = ca[n * 32 * kPitch]; ar[n] = v.x; ai[n] = v.y;
        const float2 w = cb[n * 32 * kPitch]; br[n] = w.x; bi[n] = w.y;
    }
    dft5(ar, ai, Ar, Ai);
    dft5(br, bi, Br, Bi);
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const int kp = (5 - k1) % 5;
        float2 xf, xp;
        split_pair(Ar[k1], Ai[k1], Br[kp], Bi[kp], kSplitSin[k2][k1], kSplitCos[k2][k1], xf, xp);
        float2 pf = bin_payload<MASK, BWD>(xf, vr[k1], vi[k1]);
        float2 pp = bin_payload<MASK, BWD>(xp, vr[5 + k1], vi[5 + k1]);
        if (k1 == 0) {                                    // bins 0 and 160 (real) share slot 0
            const bool z = k2 == 0;
            pf = make_float2(pf.x, z ? pp.x : pf.y);
            pp = make_float2(z ? pf.x : pp.x, z ? pf.y : pp.y);
        }
        ca[k1 * 32 * kPitch] = pf;
        cb[kp * 32 * kPitch] = pp;
    }
}

// pass 2 over all 17 columns, two at a time; the masks of the next pair are loaded into
// registers while the current pair is being computed.
template <int MASK, bool BWD>
LMFB_HD void fft_pass2_masked(float2* __restrict__ col, const float* __restrict__ mr,
                              const float* __restrict__ mi, unsigned sf) {
    float r0[10], i0[10], r1[10], i1[10];
    load_step_masks<MASK, BWD>(0, mr, mi, sf, r0, i0);
    load_step_masks<MASK, BWD>(1, mr, mi, sf, r1, i1);
#pragma unroll 1
    for (int k2 = 0; k2 < 16; k2 += 2) {
        float nr0[10], ni0[10], nr1[10], ni1[10];
        load_step_masks<MASK, BWD>(k2 + 2, mr, mi, sf, nr0, ni0);
        if (k2 + 3 <= 16) load_step_masks<MASK, BWD>(k2 + 3, mr, mi, sf, nr1, ni1);
        pass2_step<MASK, BWD>(col, k2, r0, i0);
        pass2_step<MASK, BWD>(col, k2 + 1, r1, i1);
#pragma unroll
        for (int i = 0; i < 10; ++i) { r0[i] = nr0[i]; i0[i] = ni0[i]; r1[i] = nr1[i]; i1[i] = ni1[i]; }
    }
    pass2_step<MASK, BWD>(col, 16, r0, i0);
}

// ---------------------------------------------------------------------------------------
// phase 3 (forward): banded mel accumulation filter by filter (E[m] parked in the free .y of
// slot 1+m), then log1p + store for all filters in an unrolled second sweep.
//   out : out + n*stride_n + t (row m at + m*som); inrow: t < Tmax; valid: t < T_i
// ---------------------------------------------------------------------------------------
LMFB_HD void phase3_fwd(float2* __restrict__ col, const MelBand& mb,
                        float* __restrict__ out, unsigned som, bool inrow, bool valid) {
    float* colf = reinterpret_cast<float*>(col);
    const int n_mels = mb.n_mels;
    int f = 0;
    float acc0 = 0.0f, acc1 = 0.0f;
    float* ep = colf + 2 * kPitch + 1;                    // slot 1, .y
#pragma unroll 1
    for (int m = 0; m < n_mels; ++m) {
        const int fe = mb.fend[m];
#pragma unroll 4
        for (; f < fe; ++f) {
            const float p = colf[mb.ent[f].off];
            acc0 = fmaf(mb.ent[f].wl, p, acc0);
            acc1 = fmaf(mb.ent[f].wh, p, acc1);
        }
        *ep = acc0;
        ep += 2 * kPitch;
        acc0 = acc1;
        acc1 = 0.0f;
    }
    ep = colf + 2 * kPitch + 1;
#pragma unroll 8
    for (int m = 0; m < n_mels; ++m) {
        const float y = valid ? log1pf(ep[m * 2 * kPitch]) : 0.0f;
        if (inrow) out[(unsigned)m * som] = y;
    }
}

// ---------------------------------------------------------------------------------------
// phase 3 (backward): dP[f] = wl*dE[ml] + wh*dE[ml+1]; gradients = payload * dP.
//   dE : dE + n*stride_n + t (row m at + m*sem), zero for frames t >= T_i; readable for every
//   lane (clamped by the caller).  The dE values are consumed in filter order through a sliding
//   register window: filters are handled four at a time and the window runs kDAhead filters
//   ahead; the caller pre-loads it (before the FFT) with dE[0 .. kDWin).
// ---------------------------------------------------------------------------------------
constexpr int kDAhead = 8;
constexpr int kDWin = 5 + kDAhead;                        // d[j], d[j+1] for j < 4, plus the look-ahead

LMFB_HD void dwin_preload(const float* __restrict__ dE, unsigned sem, int n_mels, float (&dw)[kDWin]) {
#pragma unroll
    for (int i = 0; i < kDWin; ++i) dw[i] = i < n_mels ? LMFB_LDG(dE + (unsigned)i * sem) : 0.0f;
}

template <int MASK>
LMFB_HD void phase3_bwd(const float2* __restrict__ col, const MelBand& mb,
                        const float* __restrict__ dE, unsigned sem, float (&dw)[kDWin],
                        float* __restrict__ gr, float* __restrict__ gi, unsigned gsf, bool inrow) {
    const int n_mels = mb.n_mels;
    int f = 0;
#pragma unroll 1
    for (int m0 = 0; m0 < n_mels; m0 += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (m0 + j < n_mels) {
                const int fe = mb.fend[m0 + j];
                const float d0 = dw[j], d1 = dw[j + 1];
#pragma unroll 2
                for (; f < fe; ++f) {
                    const uint32_t off = mb.ent[f].off, sel = mb.ent[f].sel;
                    const float dp = fmaf(mb.ent[f].wh, d1, mb.ent[f].wl * d0);
                    const float2 v = col[off >> 1];
                    const float a = sel == 2u ? v.y : v.x;
                    const float b = sel != 0u ? 0.0f : v.y;
                    if (inrow) gr[(unsigned)f * gsf] = a * dp;
                    if (MASK == kMaskReim) { if (inrow) gi[(unsigned)f * gsf] = b * dp; }
                }
            }
        }
#pragma unroll
        for (int i = 0; i + 4 < kDWin; ++i) dw[i] = dw[i + 4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + kDWin + i;
            dw[kDWin - 4 + i] = m < n_mels ? LMFB_LDG(dE + (unsigned)m * sem) : 0.0f;
        }
    }
    if (inrow) {
#pragma unroll 1
        for (; f < kBins; ++f) {                         // bins above the last filter: no gradient
            gr[(unsigned)f * gsf] = 0.0f;
            if (MASK == kMaskReim) gi[(unsigned)f * gsf] = 0.0f;
        }
    }
}

// L2 prefetch of the (32+1)*160 samples of a tile (165 lines of 128 B)
LMFB_HD void prefetch_wave_l2(int lane, const float* __restrict__ wave_row, int len, int t0) {
    long long lo = (long long)(t0 - 1) * kHop, hi = (long long)(t0 + kTile) * kHop;
    if (lo < 0) lo = 0;
    if (hi > len) hi = len;
#pragma unroll 1
    for (long long i = lo + lane * 32; i < hi; i += 32 * 32) LMFB_PREFETCH_L2(wave_row + i);
}

// L2 prefetch of the mask rows a tile will read: lanes take rows lane, lane+32, ...; a 128-byte
// row segment may straddle two lines, so both ends are touched.
LMFB_HD void prefetch_rows_l2(int lane, const float* __restrict__ base, unsigned sf, int rows, int t0, int tmax) {
    if (t0 >= tmax) return;
    const int last = (t0 + kTile <= tmax ? t0 + kTile : tmax) - 1;
#pragma unroll 1
    for (int f = lane; f < rows; f += 32) {
        LMFB_PREFETCH_L2(base + (unsigned)f * sf + t0);
        LMFB_PREFETCH_L2(base + (unsigned)f * sf + last);
    }
}

}  // namespace aas_lmfb

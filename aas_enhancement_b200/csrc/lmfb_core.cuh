// Per-thread building blocks of the fused STFT -> mask -> mel (-> log1p) kernels.
//
// Execution model ("lane = frame"): a CTA of W warps owns a tile of 32 consecutive frames of
// one utterance.  Lane t of EVERY warp works on frame t0+t, whose 320-point real FFT lives in
// column t of a shared-memory scratch; the W warps split the independent pieces of that FFT
// between them (sub-transforms in pass 1, column pairs in pass 2, filter ranges in phase 3) and
// meet at block barriers between the passes.  Consequences:
//   * every twiddle is a literal (the FFT code is identical for all lanes),
//   * every global access of a (N, F, T)/(N, M, T) tensor has the lanes along T, i.e. is a
//     128-byte coalesced row segment -- no transposition is ever needed,
//   * W times more warps are resident per SM for the same shared memory (the scratch, 42 KB per
//     32 frames, is what limits residency), which is what hides the load and FFT latencies.
//
// Real FFT of 320 samples = complex FFT of the 160 packed samples z[j] = x[2j] + i x[2j+1],
// computed with the Good-Thomas prime-factor map 160 = 5 x 32 (no inter-stage twiddles):
//   stage  : raw samples -> private columns, permuted into PFA input order (asynchronous copies)
//   pass 1 : window, then five 32-point FFTs in registers (generated codelet), in place
//   pass 2 : for each index pair (k2, 32-k2): two 5-point DFTs + the real-split butterfly give
//            ten bins; the mask(s) for those ten bins (prefetched one step ahead into
//            registers, lanes along T) are applied at once.  Forward: the masked power replaces
//            the spectrum in place (bin f lives in slot (f mod 5)*32 + (f mod 32); bins 0 and
//            160, both real, share slot 0).  Backward: dP of the ten bins is formed from dE and
//            the mask gradients are stored straight from here -- there is no phase 3.
//   phase 3: (forward only) banded mel accumulation over ascending bins out of the scratch,
//            then log1p.
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
// make a pointer opaque to the optimiser: it is then kept as one 64-bit register pair and
// "pointer + 32-bit byte offset" costs one or two instructions instead of a re-derived 64-bit
// element-offset chain (shift + add + add).
#  define LMFB_OPAQUE(p) asm volatile("" : "+l"(p))
#else
#  define LMFB_OPAQUE(p) ((void)0)
#  define LMFB_LDG(p) (*(p))
#  define LMFB_PREFETCH_L2(p) ((void)(p))
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace aas_lmfb {

constexpr int kNfft  = 320;
constexpr int kHop   = 160;
constexpr int kBins  = 161;
constexpr int kTile  = 32;             // frames per tile
constexpr int kPitch = 33;             // float2 per scratch slot (32 lanes + 1 pad: conflict-free staging)
constexpr int kSlots = 160;
constexpr int kMaxMels = 128;
constexpr int kScratchBytes = kSlots * kPitch * 8;     // 42,240 B per tile
constexpr int kMaxW = 8;               // most warps that may share a tile

enum MaskMode { kMaskNone = 0, kMaskReim = 1, kMaskPower = 2 };

// which mask tensors a kernel variant reads
#define LMFB_NEEDS_MASK_R(MASK, BWD) ((BWD) ? (MASK) == kMaskReim : (MASK) != kMaskNone)
#define LMFB_NEEDS_MASK_I(MASK, BWD) ((MASK) == kMaskReim)

// Banded-2 description of the mel basis, passed by value as a kernel parameter (constant
// bank).  Bin f feeds filters ml(f) (weight wl) and ml(f)+1 (weight wh) with ml non-decreasing,
// so the bins whose lower filter is m form the contiguous range [fend[m-1], fend[m]).
// Weights already carry the 1/4 that undoes X' = 2X.
struct BinEnt {
    float    wl, wh;   // forward: weights into filters ml, ml+1.  backward: weights of dE rows dlo, dlo+1
    uint32_t off;      // forward: float offset of the bin's masked power inside a scratch column
                       // backward: dlo * (BYTES per dE row), dlo+1 always being a valid row
    uint32_t moff;     // f * (BYTES per mask row)
};
struct MelBand {
    BinEnt  ent[kBins];
    uint8_t fend[kMaxMels];
    uint8_t mbeg[kMaxW + 1];   // forward phase 3: warp w owns filters [mbeg[w], mbeg[w+1])
    uint8_t pad_[3];
    int     n_mels;
};

// Shared-memory copy of every table the tile loop indexes at run time.  Indexed constant-bank
// loads (LDC with a register index) miss the small indexed-constant cache about one time in
// five on this kernel and each miss stalls the warp for hundreds of cycles (ncu:
// idc__request_hit_rate 79 %); shared memory has a fixed ~30-cycle latency and broadcasts.
// Filled once per (persistent) CTA by tables_fill().
struct Tables {
    float    wl[kBins + 3], wh[kBins + 3];
    uint32_t off[kBins + 3];      // forward: float offset of the bin's masked power in a scratch column
                                  // backward: byte offset of dE row dlo (dlo+1 always valid)
    float    ssin[88], scos[88];  // real-split twiddles [k2*5 + k1]
    uint8_t  binof[88];           // bin produced by pass-2 step k2, output k1 [k2*5 + k1]
    uint8_t  fend[kMaxMels];
    uint8_t  mbeg[kMaxW + 1];
    uint8_t  pad_[3];
    int      n_mels;
    uint32_t msf_bytes;           // bytes per mask row
    uint32_t pad2_[2];
};
constexpr int kTablesBytes = (int)((sizeof(Tables) + 15) / 16 * 16);
constexpr int kSmemBytes = kScratchBytes + kTablesBytes;

// Cooperative copy, once per persistent CTA.  The source lives in constant banks (kernel
// parameters and __constant__ tables), which only serve a warp quickly when all lanes read the
// SAME address: every warp therefore walks a contiguous share of the entries with a warp-uniform
// index and all lanes store the same value (a benign broadcast store).  A lane-indexed copy costs
// ~5 us per CTA (32-way serialised constant reads); this one well under 1 us.
LMFB_HD void tables_fill(Tables* tb, const MelBand& mb, uint32_t msf_bytes, int warp, int nwarps) {
    const int per = (kBins + nwarps - 1) / nwarps;
    const int f0 = warp * per, f1 = f0 + per < kBins ? f0 + per : kBins;
#pragma unroll 4
    for (int f = f0; f < f1; ++f) {
        tb->wl[f] = mb.ent[f].wl;
        tb->wh[f] = mb.ent[f].wh;
        tb->off[f] = mb.ent[f].off;
    }
    const int per2 = (85 + nwarps - 1) / nwarps;
    const int i0 = warp * per2, i1 = i0 + per2 < 85 ? i0 + per2 : 85;
    int k2 = i0 / 5, k1 = i0 - 5 * k2;
#pragma unroll 4
    for (int i = i0; i < i1; ++i) {
        tb->ssin[i] = kSplitSin[k2][k1];
        tb->scos[i] = kSplitCos[k2][k1];
        tb->binof[i] = kBinOf[k2][k1];
        if (++k1 == 5) { k1 = 0; ++k2; }
    }
    const int per3 = (kMaxMels + nwarps - 1) / nwarps;
    const int m0 = warp * per3, m1 = m0 + per3 < kMaxMels ? m0 + per3 : kMaxMels;
#pragma unroll 4
    for (int m = m0; m < m1; ++m) tb->fend[m] = mb.fend[m];
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i <= kMaxW; ++i) tb->mbeg[i] = mb.mbeg[i];
        tb->n_mels = mb.n_mels;
        tb->msf_bytes = msf_bytes;
    }
}

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

// pointer + byte offset (offsets come from the per-launch tables, in bytes, and fit 32 bits)
LMFB_HD const float* at_bytes(const float* p, uint32_t bytes) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p) + bytes);
}
LMFB_HD float* at_bytes(float* p, uint32_t bytes) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(p) + bytes);
}

#ifdef __CUDACC__
// predicated 4-byte global store (keeps the store a single predicated instruction)
__device__ __forceinline__ void st_if(float* p, float v, bool pred) {
#ifdef LMFB_DBG_NOSTORE
    if (v == 1.2345e-30f) *p = v;      // experiment: keep the value live, never store
    return;
#endif
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}"
                 :: "l"(p), "f"(v), "r"((int)pred));
}
// predicated shared-memory store that the optimiser does not see as a memory access: used for
// the E[m] slots of phase 3, which can never alias the P[f] words the same loop reads, so that
// those reads may be hoisted and overlapped freely.
__device__ __forceinline__ void sts_if_noalias(float* p, float v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.shared.f32 [%0], %1;\n\t}"
                 :: "r"((unsigned)__cvta_generic_to_shared(p)), "f"(v), "r"((int)pred));
}
#else
static inline void st_if(float* p, float v, bool pred) { if (pred) *p = v; }
static inline void sts_if_noalias(float* p, float v, bool pred) { if (pred) *p = v; }
#endif

// ---------------------------------------------------------------------------------------
// Staging: the (32+1)*160 samples a tile needs go STRAIGHT from global memory into the 32 frame
// columns with 8-byte asynchronous copies (cp.async / LDGSTS): no register round trip, every
// copy of the tile in flight at once, one exposed memory round trip per tile instead of one per
// load batch.  The samples land RAW and in PFA input order; the window is applied by pass 1 when
// it loads them (the window table lives in the pad column of the scratch, window_fill()).
// Lanes run along the packed-sample index of a hop-row, so the global side is a coalesced
// 256-byte run and the shared side is conflict-free (slot pitch 33).  Hop-row r feeds frame r
// (first half, r < 32) and frame r-1 (second half, r >= 1): two copies of the same 8 bytes, the
// second of which hits L1.
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void cp_async8(float2* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"
                 :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}
#else
static inline void cp_async8(float2* dst, const float* src) { dst->x = src[0]; dst->y = src[1]; }
static inline void cp_async_wait_all() {}
#endif

struct StageLane {                      // per-lane constants of the staging map (float2 index of slot * pitch)
    int slot_a[3], slot_b[3];
};

LMFB_HD void stage_lane_init(int lane, StageLane& sl) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int c = lane + 32 * q;                       // packed index inside a hop-row, < 80 valid
        const int cc = c < 80 ? c : 0;
        sl.slot_a[q] = slot_of_packed(cc) * kPitch;
        sl.slot_b[q] = slot_of_packed(cc + 80) * kPitch;
    }
}

// window table: pad column (index 32) of slot s holds the window pair of the packed sample that
// lives in slot s.  Written once per persistent CTA; staging and the passes never touch column 32.
LMFB_HD void window_fill(float2* __restrict__ S, const float* __restrict__ window, int idx, int cnt) {
    for (int j = idx; j < kSlots; j += cnt)
        S[slot_of_packed(j) * kPitch + kTile] = make_float2(LMFB_LDG(window + 2 * j), LMFB_LDG(window + 2 * j + 1));
}

// hop-row q of the unpadded signal lies fully inside [0, len) and can be copied as 8-byte pieces
LMFB_HD bool row_interior(int q, int len, bool vec_ok) {
    return vec_ok && q >= 0 && (long long)(q + 1) * kHop <= (long long)len;
}

LMFB_HD void stage_store(const StageLane& sl, float2* __restrict__ S, int r, int q, float2 v) {
    if (r < kTile)  S[sl.slot_a[q] + r]     = v;
    if (r >= 1)     S[sl.slot_b[q] + r - 1] = v;
}

// Only hop-rows 0 .. n_rows-1 feed a frame that exists (n_rows = valid frames of the tile + 1);
// the others are zero-filled without touching global memory.  Rows that need the reflect padding
// (either end of the utterance) or an unaligned wave take a per-sample path with plain loads.
// The 33 rows are dealt to the W warps in contiguous shares; every branch is warp-uniform.
template <int W>
LMFB_HD void stage_tile(int w, int lane, const StageLane& sl, const float* __restrict__ wave_row, int len,
                        int t0, int n_rows, float2* __restrict__ S, bool vec_ok) {
    constexpr int kShare = (kTile + 1 + W - 1) / W;               // rows per warp (the last warp may have fewer)
    const int r_lo = w * kShare;
    const int r_hi = r_lo + kShare < kTile + 1 ? r_lo + kShare : kTile + 1;
#pragma unroll 1
    for (int r = r_lo; r < r_hi; ++r) {
        const int q = t0 + r - 1;                                 // hop-row of the signal
        if (r >= n_rows) {
            const float2 z = make_float2(0.0f, 0.0f);
            stage_store(sl, S, r, 0, z);
            stage_store(sl, S, r, 1, z);
            if (lane < 16) stage_store(sl, S, r, 2, z);
        } else if (row_interior(q, len, vec_ok)) {
            const float* src = wave_row + (long long)q * kHop + 2 * lane;
            if (r < kTile) {
                cp_async8(S + sl.slot_a[0] + r, src);
                cp_async8(S + sl.slot_a[1] + r, src + 64);
                if (lane < 16) cp_async8(S + sl.slot_a[2] + r, src + 128);
            }
            if (r >= 1) {
                cp_async8(S + sl.slot_b[0] + r - 1, src);
                cp_async8(S + sl.slot_b[1] + r - 1, src + 64);
                if (lane < 16) cp_async8(S + sl.slot_b[2] + r - 1, src + 128);
            }
        } else {
            const int base = q * kHop;
            float2 v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {                         // all six loads in flight before the stores
                const int c = lane + 32 * k;
                v[k] = make_float2(0.0f, 0.0f);
                if (c < 80) {
                    v[k].x = LMFB_LDG(wave_row + reflect_index(base + 2 * c, len));
                    v[k].y = LMFB_LDG(wave_row + reflect_index(base + 2 * c + 1, len));
                }
            }
            stage_store(sl, S, r, 0, v[0]);
            stage_store(sl, S, r, 1, v[1]);
            if (lane < 16) stage_store(sl, S, r, 2, v[2]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// pass 1: the five in-register 32-point FFTs of a column, dealt round-robin to the W warps
// ---------------------------------------------------------------------------------------
template <int W>
LMFB_HD void fft_pass1(int w, float2* __restrict__ col, const float2* __restrict__ win) {
#pragma unroll 1
    for (int n1 = w; n1 < 5; n1 += W) {
        float2* p = col + n1 * 32 * kPitch;
        const float2* wn = win + n1 * 32 * kPitch;                // same address for every lane: broadcast
        float xr[32], xi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float2 v = p[i * kPitch], g = wn[i * kPitch];
            xr[i] = v.x * g.x; xi[i] = v.y * g.y;
        }
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

// forward: masked power P' of one bin
template <int MASK>
LMFB_HD float masked_power(float2 x, float mr, float mi) {
    if (MASK == kMaskReim) { const float a = x.x * mr, b = x.y * mi; return fmaf(a, a, b * b); }
    const float p = fmaf(x.x, x.x, x.y * x.y);
    return MASK == kMaskPower ? mr * p : p;
}

// what pass 2 needs from global memory for the ten bins of column pair k2: index k1 -> bin f,
// index 5+k1 -> bin 160-f.  All pointers are readable for every lane (out-of-row lanes are
// clamped by the caller), so no load is predicated; offsets come from the per-launch tables.
struct StepMasks { float vr[10], vi[10]; };              // mask values (prefetched one step ahead)
struct StepD     { float d0[10], d1[10]; };              // backward: the two dE rows of each bin

template <int MASK, bool BWD>
LMFB_HD void load_masks(int k2, const Tables& tb, const float* __restrict__ mr, const float* __restrict__ mi,
                        StepMasks& in) {
    const uint32_t msf = tb.msf_bytes;
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const unsigned f = tb.binof[k2 * 5 + k1], fp = kBins - 1 - f;
        const uint32_t of = f * msf, op = fp * msf;
#ifdef LMFB_DBG_NOMASKLOAD
        in.vr[k1] = in.vr[5 + k1] = 0.5f + 1e-9f * of; in.vi[k1] = in.vi[5 + k1] = 0.25f + 1e-9f * op;   // experiment
        continue;
#endif
        if (LMFB_NEEDS_MASK_R(MASK, BWD)) { in.vr[k1] = LMFB_LDG(at_bytes(mr, of)); in.vr[5 + k1] = LMFB_LDG(at_bytes(mr, op)); }
        if (LMFB_NEEDS_MASK_I(MASK, BWD)) { in.vi[k1] = LMFB_LDG(at_bytes(mi, of)); in.vi[5 + k1] = LMFB_LDG(at_bytes(mi, op)); }
    }
}

// issued at the start of the step that consumes it: the two 5-point DFTs and the split (about
// 200 instructions that need nothing from global memory) run while these loads are in flight.
// `dE` is either the global column (row stride sem_bytes) or, in the variants that stage the
// tile's dE rows in shared memory, that staged column (row stride 128 bytes).
template <bool DSMEM>
LMFB_HD void load_d(int k2, const Tables& tb, const float* __restrict__ dE, unsigned sem_bytes, StepD& in) {
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const unsigned f = tb.binof[k2 * 5 + k1], fp = kBins - 1 - f;
        const uint32_t df = tb.off[f], dp = tb.off[fp];
        if (DSMEM) {                                           // shared memory: plain loads
            in.d0[k1]     = *at_bytes(dE, df);
            in.d1[k1]     = *at_bytes(dE, df + sem_bytes);
            in.d0[5 + k1] = *at_bytes(dE, dp);
            in.d1[5 + k1] = *at_bytes(dE, dp + sem_bytes);
        } else {
            in.d0[k1]     = LMFB_LDG(at_bytes(dE, df));
            in.d1[k1]     = LMFB_LDG(at_bytes(dE, df + sem_bytes));
            in.d0[5 + k1] = LMFB_LDG(at_bytes(dE, dp));
            in.d1[5 + k1] = LMFB_LDG(at_bytes(dE, dp + sem_bytes));
        }
    }
}

// ---------------------------------------------------------------------------------------
// pass 2, one step: columns k2 and kb = (32-k2) mod 32 of the five sub-transforms -> two 5-point
// DFTs -> real split -> ten bins -> mask.  Branch-free and identical for all k2: the self-paired
// columns (k2 = 0, 16, where kb == k2) simply compute each of their bins twice.
//   forward : masked power stored in place (bins 0/160 share slot 0 -> one select)
//   backward: gradients = (2 Mr Re'^2, 2 Mi Im'^2) * dP, or (Re'^2 + Im'^2) * dP for 'power',
//             stored to gr/gi (+ f*gsf); every lane's pointers are valid, `inrow` gates the store
// ---------------------------------------------------------------------------------------
template <int MASK, bool BWD, bool DSMEM>
LMFB_HD void pass2_step(float2* __restrict__ col, int k2, const Tables& tb, const StepMasks& in,
                        const float* __restrict__ dE, unsigned sem_bytes,
                        float* __restrict__ gr, float* __restrict__ gi, bool inrow) {
    StepD d;
    if (BWD) load_d<DSMEM>(k2, tb, dE, sem_bytes, d);
    const int kb = (32 - k2) & 31;
    float2* ca = col + k2 * kPitch;
    float2* cb = col + kb * kPitch;
    float ar[5], ai[5], br[5], bi[5], Ar[5], Ai[5], Br[5], Bi[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) {
        const float2 v = ca[n * 32 * kPitch]; ar[n] = v.x; ai[n] = v.y;
        const float2 u = cb[n * 32 * kPitch]; br[n] = u.x; bi[n] = u.y;
    }
    dft5(ar, ai, Ar, Ai);
    dft5(br, bi, Br, Bi);
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const int kp = (5 - k1) % 5;
        float2 xf, xp;
        split_pair(Ar[k1], Ai[k1], Br[kp], Bi[kp], tb.ssin[k2 * 5 + k1], tb.scos[k2 * 5 + k1], xf, xp);
        if (!BWD) {
            float pf = masked_power<MASK>(xf, in.vr[k1], in.vi[k1]);
            float pp = masked_power<MASK>(xp, in.vr[5 + k1], in.vi[5 + k1]);
            float2 sf2 = make_float2(pf, 0.0f), sp2 = make_float2(pp, 0.0f);
            if (k1 == 0) {                                // bins 0 and 160 (real) share slot 0
                const bool z = k2 == 0;
                sf2 = make_float2(pf, z ? pp : 0.0f);
                sp2 = make_float2(z ? pf : pp, z ? pp : 0.0f);
            }
            ca[k1 * 32 * kPitch] = sf2;
            cb[kp * 32 * kPitch] = sp2;
        } else {
            const unsigned f = tb.binof[k2 * 5 + k1], fp = kBins - 1 - f;
            const uint32_t of = f * tb.msf_bytes, op = fp * tb.msf_bytes;
            const float dpf = fmaf(tb.wh[f], d.d1[k1], tb.wl[f] * d.d0[k1]);
            const float dpp = fmaf(tb.wh[fp], d.d1[5 + k1], tb.wl[fp] * d.d0[5 + k1]);
            if (MASK == kMaskReim) {
                st_if(at_bytes(gr, of), 2.0f * in.vr[k1] * xf.x * xf.x * dpf, inrow);
                st_if(at_bytes(gi, of), 2.0f * in.vi[k1] * xf.y * xf.y * dpf, inrow);
                st_if(at_bytes(gr, op), 2.0f * in.vr[5 + k1] * xp.x * xp.x * dpp, inrow);
                st_if(at_bytes(gi, op), 2.0f * in.vi[5 + k1] * xp.y * xp.y * dpp, inrow);
            } else {
                st_if(at_bytes(gr, of), fmaf(xf.x, xf.x, xf.y * xf.y) * dpf, inrow);
                st_if(at_bytes(gr, op), fmaf(xp.x, xp.x, xp.y * xp.y) * dpp, inrow);
            }
        }
    }
}

// pass 2 over the 17 column pairs, dealt round-robin to the W warps; the global inputs of a
// warp's next step are loaded into a second register set while the current step is computed.
// `a` must already hold the masks of the warp's first step (k2 = w): the caller issues that
// load before the block barrier that ends pass 1, so its latency hides behind the barrier.
template <int W, int MASK, bool BWD, bool DSMEM>
LMFB_HD void fft_pass2(int w, float2* __restrict__ col, const Tables& tb, StepMasks& a,
                       const float* __restrict__ mr, const float* __restrict__ mi,
                       const float* __restrict__ dE, unsigned sem_bytes,
                       float* __restrict__ gr, float* __restrict__ gi, bool inrow) {
    StepMasks b;
#pragma unroll 1
    for (int k2 = w; k2 <= 16; k2 += 2 * W) {
        const bool has_b = k2 + W <= 16;
        if (has_b) load_masks<MASK, BWD>(k2 + W, tb, mr, mi, b);
        pass2_step<MASK, BWD, DSMEM>(col, k2, tb, a, dE, sem_bytes, gr, gi, inrow);
        if (has_b) {
            if (k2 + 2 * W <= 16) load_masks<MASK, BWD>(k2 + 2 * W, tb, mr, mi, a);
            pass2_step<MASK, BWD, DSMEM>(col, k2 + W, tb, b, dE, sem_bytes, gr, gi, inrow);
        }
    }
}

// ---------------------------------------------------------------------------------------
// phase 3 (forward): banded mel accumulation filter by filter (E[m] parked in the free .y of
// slot 1+m), then log1p + store in an unrolled second sweep.  Warp w owns filters
// [mbeg[w], mbeg[w+1]); to get the upper-weight contributions of its first filter it starts
// one filter early and discards that filter's (partial) sum.
//   out : out + n*stride_n + t (row m at + m*som); inrow: t < Tmax; valid: t < T_i
// ---------------------------------------------------------------------------------------
LMFB_HD void phase3_fwd(int w, float2* __restrict__ col, const Tables& tb,
                        float* __restrict__ out, unsigned som_bytes, bool inrow, bool valid
#ifdef LMFB_TIMELINE
                        , long long* g_tl_mid = nullptr
#endif
                        ) {
    float* colf = reinterpret_cast<float*>(col);
    const int m_lo = tb.mbeg[w], m_hi = tb.mbeg[w + 1];
    if (m_lo >= m_hi) return;
    const int m_first = m_lo > 0 ? m_lo - 1 : 0;
    int f = m_first > 0 ? (int)tb.fend[m_first - 1] : 0;
    float acc0 = 0.0f, acc1 = 0.0f;
    float* ep = colf + (1 + m_first) * 2 * kPitch + 1;     // E[m] -> .y of slot 1+m
    int fe = tb.fend[m_first];
#pragma unroll 1
    for (int m = m_first; m < m_hi; ++m) {
        const int fe_next = m + 1 < m_hi ? (int)tb.fend[m + 1] : fe;     // fetched one filter ahead
#pragma unroll 4
        for (; f < fe; ++f) {
            // slot of bin f computed, not looked up: the P read does not wait for a table read
            const int off = ((f % 5) * 32 + (f & 31)) * (2 * kPitch) + (f == kBins - 1 ? 1 : 0);
            const float p = colf[off];
            acc0 = fmaf(tb.wl[f], p, acc0);
            acc1 = fmaf(tb.wh[f], p, acc1);
        }
        if (m >= m_lo) *ep = acc0;                         // the early filter m_lo-1 belongs to another warp
        ep += 2 * kPitch;
        acc0 = acc1;
        acc1 = 0.0f;
        fe = fe_next;
    }
#ifdef LMFB_TIMELINE
    if (g_tl_mid) *g_tl_mid = clock64();
#endif
    const float* eq = colf + (1 + m_lo) * 2 * kPitch + 1;
    float* op = at_bytes(out, (uint32_t)m_lo * som_bytes);
#pragma unroll 4
    for (int m = m_lo; m < m_hi; ++m) {
        const float y = valid ? log1pf(*eq) : 0.0f;
        st_if(op, y, inrow);
        eq += 2 * kPitch;
        op = at_bytes(op, som_bytes);
    }
}

// L2 prefetch of the (32+1)*160 samples of a tile (165 lines of 128 B), `idx`/`cnt` = this
// thread's index / the number of threads sharing the job
LMFB_HD void prefetch_wave_l2(int idx, int cnt, const float* __restrict__ wave_row, int len, int t0) {
    long long lo = (long long)(t0 - 1) * kHop, hi = (long long)(t0 + kTile) * kHop;
    if (lo < 0) lo = 0;
    if (hi > len) hi = len;
#pragma unroll 1
    for (long long i = lo + idx * 32; i < hi; i += (long long)cnt * 32) LMFB_PREFETCH_L2(wave_row + i);
}

// L2 prefetch of the row segments a tile will read: threads take rows idx, idx+cnt, ...; a
// 128-byte segment may straddle two lines, so both ends are touched.
LMFB_HD void prefetch_rows_l2(int idx, int cnt, const float* __restrict__ base, unsigned sf, int rows, int t0, int tmax) {
    if (t0 >= tmax) return;
    const int last = (t0 + kTile <= tmax ? t0 + kTile : tmax) - 1;
#pragma unroll 1
    for (int f = idx; f < rows; f += cnt) {
        LMFB_PREFETCH_L2(base + (unsigned)f * sf + t0);
        LMFB_PREFETCH_L2(base + (unsigned)f * sf + last);
    }
}

}  // namespace aas_lmfb

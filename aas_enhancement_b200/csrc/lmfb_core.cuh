// Per-thread building blocks of the fused STFT -> mask -> mel (-> log1p) kernels.
//
// Execution model ("lane = frame"): a CTA of W warps owns a tile of 32 consecutive frames of
// one utterance.  Lane t of EVERY warp works on frame t0+t, whose 320-point real FFT lives in
// column t of a shared-memory scratch; the W warps split the independent pieces of that FFT
// between them (sub-transforms in pass 1, column pairs in pass 2, bin ranges in phase 3) and
// meet at block barriers between the passes.  Consequences:
//   * every twiddle is a literal (the FFT code is identical for all lanes),
//   * every global access of a (N, F, T)/(N, M, T) tensor has the lanes along T, i.e. is a
//     128-byte coalesced row segment -- no transposition is ever needed,
//   * W times more warps are resident per SM for the same shared memory (the scratch, 42 KB per
//     32 frames, is what limits residency).
//
// Real FFT of 320 samples = complex FFT of the 160 packed samples z[j] = x[2j] + i x[2j+1],
// computed with the Good-Thomas prime-factor map 160 = 5 x 32 (no inter-stage twiddles):
//   stage  : raw samples -> private columns, permuted into PFA input order (asynchronous copies)
//   pass 1 : window, then five 32-point FFTs in registers (generated codelet), in place
//   pass 2 : for each index pair (k2, 32-k2): two 5-point DFTs + the real-split butterfly give
//            ten bins; the mask(s) for those ten bins (prefetched one step ahead into
//            registers, lanes along T) are applied at once.  Forward: the masked power P[f] is
//            written back over the step's own slots as a compact float row.  Backward: dP of the
//            ten bins is formed from dE and the mask gradients are stored straight from here.
//   phase 3: (forward only) banded mel accumulation over ascending bins, then log1p.
// The code is rolled (one copy of a pass-2 step and of an 8-bin phase-3 group serves every warp
// and stays in the 32 KB instruction cache); what differs between steps / bins comes from small
// tables in shared memory, laid out so that a step fetches them with a few 128-bit broadcast
// loads, and every global row address is ONE IMAD.WIDE (row * stride + base).  Measured
// alternatives: per-element tables with 32-bit offsets (16.9 k / 17.4 k warp instructions per
// tile forward / backward, 60 % of them address arithmetic and table loads) and fully unrolled
// passes with compile-time bins (fewest instructions, but 145 KB of code per kernel: 15 warps
// streaming different code thrash the instruction cache, twice slower).
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
#include <type_traits>
#include <string.h>
#include "fft_codelets.cuh"

#ifdef __CUDACC__
#  define LMFB_CX __host__ __device__ constexpr
#else
#  define LMFB_CX constexpr
#endif

#ifdef __CUDACC__
#  define LMFB_LDG(p) __ldg(p)
#  define LMFB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
// make a value opaque to the optimiser: it is then kept in its register(s) instead of being
// re-derived (rematerialised) at every use
#  define LMFB_OPAQUE(p) asm volatile("" : "+l"(p))
#  define LMFB_OPAQUE32(v) asm volatile("" : "+r"(v))
#  define LMFB_SYNCWARP() __syncwarp()
#else
#  define LMFB_SYNCWARP() ((void)0)
#  define LMFB_OPAQUE(p) ((void)0)
#  define LMFB_OPAQUE32(v) ((void)0)
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
constexpr int kRow   = 2 * kPitch;     // floats per scratch slot row
constexpr int kSlots = 160;
constexpr int kMaxMels = 128;
constexpr int kScratchBytes = kSlots * kPitch * 8 + 128;   // 42,240 B of slots + one 32-float row for the power of bin 160
constexpr int kMaxW = 8;               // most warps that may share a tile

enum MaskMode { kMaskNone = 0, kMaskReim = 1, kMaskPower = 2,
                kStftOut = 3 };   // "backward" skeleton that stores the spectrum itself (Re rows to gr, Im rows to gi)

// which mask tensors a kernel variant reads
// (the backward needs the 'power' mask only for the gradient into the waveform, GW)
#define LMFB_NEEDS_MASK_R(MASK, BWD, GW) ((BWD) ? ((MASK) == kMaskReim || ((GW) && (MASK) == kMaskPower)) : ((MASK) != kMaskNone && (MASK) != kStftOut))
#define LMFB_NEEDS_MASK_I(MASK, BWD, GW) ((MASK) == kMaskReim)

// compile-time loop: f(IC<B>{}), f(IC<B+1>{}), ... f(IC<E-1>{})
template <int I> struct IC { static constexpr int value = I; };
template <int B, int E, class F>
LMFB_HD void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(static_cast<F&&>(f));
    }
}
// run f(IC<w>{}) for the (warp-uniform) run-time w in [0, W)
template <int W, class F>
LMFB_HD void dispatch_warp(int w, F&& f) {
    static_for<0, W>([&](auto wc) { if (w == decltype(wc)::value) f(wc); });
}

// ---- tables -------------------------------------------------------------------------------
// Forward: "walkable" bases (banded-2, see FwdTab) are walked bin by bin with two running sums;
// any other basis (dense, re-ordered, hand-made) is a GATHER per filter over its row [lo, lo + cnt)
// with the weights read from the caller's device copy of the (M, 161) matrix.  The spectrum is
// carried as X' = 2X; the 1/4 is folded into the walk's weights / applied at the end of the gather.
// Backward: per pass-2 step and output, the two dE rows (d, d+1) a bin's gradient is formed from
// and their weights ("banded": every bin feeds at most two adjacent filters; other bases go
// through a precomputed dP = B^T dE, see lmfb_kernels.cu).
// The tables travel as a kernel parameter and are copied once per persistent CTA into shared
// memory behind the scratch.
constexpr int kGroups = kSlots / 8;            // phase 3 (walk) goes over bins 0..159 in groups of 8; bin 160 is peeled

struct FwdTab {                                // forward kernel parameter
    // "walkable" bases (banded-2: bin f feeds filters ml(f) (weight wl) and ml(f)+1 (weight wh) with ml
    // non-decreasing -- every triangular filterbank); weights carry the 1/4 that undoes X' = 2X
    float2   w[kBins];                         // (wl, wh) of bin f
    uint8_t  hmask[kGroups + 8];               // bit f & 7 of byte f >> 3: the band moves on before bin f (adv != 0); read as
                                               // 32-bit words at any bit offset, hence the zero word behind bin 160
    uint8_t  adv[kBins + 3];                   // ml(f) - ml(f-1): filters completed before bin f (0 for f = 0)
    // phase 3: warp w OWNS filters lo[w] .. hi[w] (lo > hi: none) and walks every bin that feeds one of
    // them, [b0[w], b1[w]): the bins whose lower filter is lo - 1 .. hi.  The walk delivers a sum for every
    // filter m0[w] .. m1[w] = ml(b0) .. ml(b1 - 1) + 1 (m0 > m1: none); those of lo - 1 and hi + 1 are
    // incomplete and belong to the neighbours.
    uint8_t  lo[kMaxW], hi[kMaxW];
    uint8_t  b0[kMaxW], b1[kMaxW];
    uint8_t  m0[kMaxW], m1[kMaxW];
    // any other basis: filter m is the row [lo, lo + cnt) of the caller's device matrix
    uint32_t row[kMaxMels];                    // lo | cnt << 8
    int      n_mels;
    int      multi;                            // some bin completes more than one filter (adv > 1)
    int      walkable;
};
struct BwdTab {                                // backward kernel parameter, per pass-2 step k2 and output k1
    float    w[17][5][4];                      // wl_f, wh_f, wl_fp, wh_fp: weights of the dE rows (d, d+1)
    uint32_t d[17][5][2];                      // dE row d of bin f and of bin fp = 160 - f (d+1 is always valid)
    int      n_mels;                           // rows of the dE tensor
    int      pad_;
};
// shared-memory image: the per-step constants of the algorithm, then the parameter above
struct alignas(16) StepEnt { uint32_t f[5]; float sn[5]; float cs[5]; uint32_t pad_; };     // 64 B
struct alignas(16) FwdSmem { StepEnt step[17]; float2 w[kBins]; uint8_t hmask[kGroups + 8]; uint8_t adv[kBins + 3]; uint32_t row[kMaxMels]; };
struct alignas(16) BwdSmem { StepEnt step[17]; float w[17][5][4]; uint32_t d[17][5][2]; };
constexpr int kTabBytesFwd = (int)((sizeof(FwdSmem) + 15) / 16 * 16);
constexpr int kTabBytesBwd = (int)((sizeof(BwdSmem) + 15) / 16 * 16);
template <bool BWD> struct TabOf;
template <> struct TabOf<false> { typedef FwdTab Param; typedef FwdSmem Smem; };
template <> struct TabOf<true>  { typedef BwdTab Param; typedef BwdSmem Smem; };
template <bool BWD> struct TabBytes { static constexpr int value = BWD ? kTabBytesBwd : kTabBytesFwd; };
// scratch | tables | control block: scheduler answer (16), its barrier (8), the raw buffer's barrier (8),
// the table copy's barrier (8), pad (8) | raw buffer
constexpr int kCtlBytes = 48;
constexpr int kRawBytes_ = (kTile + 1) * 164 * 4;
// backward: + the tile's dE rows (up to kDeSmemRows mel rows x 32 frames), see stage_de()
constexpr int kDeSmemRows = 48;
constexpr int kDeSmemBytes = kDeSmemRows * 32 * 4;
constexpr int smem_bytes(bool bwd) { return kScratchBytes + (bwd ? kTabBytesBwd : kTabBytesFwd) + kCtlBytes + kRawBytes_ + (bwd ? kDeSmemBytes : 0); }
// the device-resident image of both shared-memory tables (aas_lmfb_plan_upload): forward, then backward
constexpr int kTabBlobBytes = kTabBytesFwd + kTabBytesBwd;

LMFB_CX int slot_of_packed(int j) {            // j in [0,160): packed-sample index -> PFA input slot
    return ((3 * (j % 5)) % 5) * 32 + ((13 * (j & 31)) & 31);
}
// bin produced by pass-2 step k2, output k1 (CRT of k1 mod 5, k2 mod 32); its partner is 160 - f
LMFB_CX int bin_of(int k2, int k1) { return (96 * k1 + 65 * k2) % 160; }
// float offset (inside the scratch, before adding the lane) of the masked power of bin f after
// pass 2: the first 32 floats of slot row f (row f = columns f mod 32 of sub-transform f / 32 is
// one of the rows the producing step has just consumed); bin 160 has a row of its own behind the slots
LMFB_CX int p_off(int f) { return f * kRow; }
// float offset of sum row r of phase 3 (walk): second 32 floats of slot row 1 + r
LMFB_CX int e_off(int r) { return (1 + r) * kRow + kTile; }

// 'reflect' padding index (numpy semantics, any offset, length >= 1)
LMFB_HD int reflect_index(int i, int len) {
    if (len <= 1) return 0;
    const int period = 2 * (len - 1);
    int j = i % period;
    if (j < 0) j += period;
    return j >= len ? period - j : j;
}

// row `row` of a tensor whose rows are `stride_bytes` apart.  (Forcing one IMAD.WIDE per address with
// inline PTX -- the compiler shares the product row * stride between tensors with equal strides and
// then pays two IADD3 per address -- was measured: 700 fewer instructions per backward tile and 5 %
// SLOWER.  IMAD.WIDE issues on the FMA pipe, which the FFT already saturates; the IADD3 pairs ride the
// otherwise idle ALU pipe.)
LMFB_HD const float* at_row(const float* p, uint32_t row, uint32_t stride_bytes) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p) +
                                          (unsigned long long)row * (unsigned long long)stride_bytes);
}
LMFB_HD float* at_row(float* p, uint32_t row, uint32_t stride_bytes) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(p) +
                                    (unsigned long long)row * (unsigned long long)stride_bytes);
}

#ifdef __CUDACC__
// predicated 4-byte global store (keeps the store a single predicated instruction)
__device__ __forceinline__ void st_if(float* p, float v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}"
                 :: "l"(p), "f"(v), "r"((int)pred));
}
// the same with the streaming (evict-first) policy: the mask gradients, 330 MB per launch that
// nothing reads before the enhancer's backward, should not push the mask rows still to be read out
// of L2 (measured: backward 237 -> 234 us)
__device__ __forceinline__ void st_if_cs(float* p, float v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.f32 [%0], %1;\n\t}"
                 :: "l"(p), "f"(v), "r"((int)pred));
}
// predicated shared-memory store that the optimiser does not see as a memory access: used for
// the partial-sum rows of phase 3, which can never alias the P[f] words the same code reads, so
// that those reads may be hoisted and overlapped freely
__device__ __forceinline__ void sts_if_noalias(float* p, float v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.shared.f32 [%0], %1;\n\t}"
                 :: "r"((unsigned)__cvta_generic_to_shared(p)), "f"(v), "r"((int)pred));
}
#else
static inline void st_if(float* p, float v, bool pred) { if (pred) *p = v; }
static inline void st_if_cs(float* p, float v, bool pred) { if (pred) *p = v; }
static inline void sts_if_noalias(float* p, float v, bool pred) { if (pred) *p = v; }
#endif

// ---------------------------------------------------------------------------------------
// Staging.  The (32+1)*160 samples a tile needs are ONE contiguous, 16-byte aligned run of the
// waveform.  They land RAW in a buffer of their own, one hop-row (160 samples) per asynchronous
// bulk copy (cp.async.bulk, the TMA engine: one instruction per 640 bytes issued by one lane, no
// load/store-unit work, completion on an mbarrier), rows 164 floats apart so that the lanes of
// pass 1 (lane = frame, i.e. one row apart) spread over the banks.  Because the buffer is free
// again as soon as pass 1 has moved the tile into the FFT scratch, the NEXT tile's rows are
// requested right there and fly under pass 2 / phase 3 of the current one: no staging latency is
// exposed in steady state, and every hop-row is fetched once (the first half of frame r and the
// second half of frame r-1 are the same row).  History: 8-byte cp.async (LDGSTS) straight into the
// frame columns, every row twice -- 5,280 load/store-unit elements and 6-8 k exposed cycles per tile.
// Rows that need the reflect padding (both ends of an utterance), rows past the last frame (zeros)
// and waves that are not 16-byte aligned take a per-row path with plain loads and stores.
// ---------------------------------------------------------------------------------------
constexpr int kRawPitch = 164;                              // floats between hop-rows of the raw buffer
constexpr int kRawBytes = (kTile + 1) * kRawPitch * 4;      // 21,648 B
static_assert(kRawBytes == kRawBytes_, "raw buffer size");

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_row(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#else
static inline void bulk_row(void* dst, const void* src, unsigned bytes, uint64_t*) { memcpy(dst, src, bytes); }
static inline void mbar_arrive_tx(uint64_t*, unsigned) {}
#endif

#ifdef __CUDACC__
// backward: the tile's dE rows (n_mels x 32 frames, 5 KB for 40 mels) straight into shared memory with 4-byte
// asynchronous copies; pass 2 re-reads every row eight times, and with three 67 KB tiles per SM only 26 KB
// of L1 are left to catch those re-reads (measured on 256 x 10 s: backward 0.2053 -> 0.2021 ms).
// Rows w, w + W, ...; `de` = this lane's (clamped) column of row 0.
template <int W>
__device__ __forceinline__ void stage_de(int w, int lane, const float* __restrict__ de, unsigned sem_bytes, int n_mels,
                                         float* __restrict__ de_s) {
#pragma unroll 4
    for (int m = w; m < n_mels; m += W)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(de_s + m * 32 + lane)),
                     "l"(at_row(de, (uint32_t)m, sem_bytes)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

struct StageLane {                      // per-lane constants of the slot map (used by the adjoint staging)
    int      slot_a[3], slot_b[3];      // float2 index of (slot * pitch) for packed samples c, c + 80
};

LMFB_HD void stage_lane_init(int lane, StageLane& sl) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int c = lane + 32 * q;                       // packed index inside a hop-row, < 80 valid
        const int cc = c < 80 ? c : 0;
        sl.slot_a[q] = slot_of_packed(cc) * kPitch;
        sl.slot_b[q] = slot_of_packed(cc + 80) * kPitch;
        LMFB_OPAQUE32(sl.slot_a[q]); LMFB_OPAQUE32(sl.slot_b[q]);
    }
}

// window table: pad column (index 32) of slot s holds the window pair of the packed sample that
// lives in slot s.  Written once per persistent CTA; the passes never touch column 32.
LMFB_HD void window_fill(float2* __restrict__ S, const float* __restrict__ window, int idx, int cnt, float scale = 1.0f) {
    for (int j = idx; j < kSlots; j += cnt)
        S[slot_of_packed(j) * kPitch + kTile] = make_float2(scale * LMFB_LDG(window + 2 * j), scale * LMFB_LDG(window + 2 * j + 1));
}

// hop-row q of the unpadded signal lies fully inside [0, len) and can be copied as a whole
LMFB_HD bool row_interior(int q, int len, bool vec_ok) {
    return vec_ok && q >= 0 && (long long)(q + 1) * kHop <= (long long)len;
}

// Warp w requests / fills its share of the tile's 33 hop-rows (rows w, w + W, ...) and then signals
// the buffer's mbarrier once, with the bytes its bulk copies will deliver.  Plain stores of the
// per-row path are ordered before the consumer by the block barrier that always lies between this
// call and the pass 1 that reads the buffer.
//   I16: the wave is int16 PCM (bytes on the wire and in HBM halved): the rows land
//        as they are (320 bytes each, same row pitch) and pass 1 converts; 1/32768 is in the window table.
template <int W, bool I16>
LMFB_HD void stage_raw(int w, int lane, const void* __restrict__ wave_row_, int len, int t0, int n_rows,
                       float* __restrict__ raw, uint64_t* bar, bool vec16) {
    typedef typename std::conditional<I16, int16_t, float>::type Sample;
    const Sample* wave_row = static_cast<const Sample*>(wave_row_);
    constexpr unsigned kRowBytes = kHop * sizeof(Sample);
    // fast path (all tiles but the first and last of an utterance): every row is a whole interior hop-row
    if (vec16 && n_rows == kTile + 1 && t0 >= 1 && (t0 + kTile) * kHop <= len) {       // (t0 < 2^22: no overflow)
        if (lane == 0) {
            const Sample* src = wave_row + (long long)(t0 - 1 + w) * kHop;
            float* dst = raw + w * kRawPitch;
            // rows w, w + W, ...: kAll of them in every warp, one more in the first kMore warps (spelled out: left
            // as `if (w + i * W < 33)` the tests fold, or not, with the optimiser's knowledge of the warp index)
            constexpr int kAll = (kTile + 1) / W, kMore = (kTile + 1) % W;
#pragma unroll
            for (int i = 0; i < kAll; ++i) bulk_row(dst + i * W * kRawPitch, src + i * W * kHop, kRowBytes, bar);
            const bool more = w < kMore;
            if (more) bulk_row(dst + kAll * W * kRawPitch, src + kAll * W * kHop, kRowBytes, bar);
            mbar_arrive_tx(bar, (unsigned)((kAll + (more ? 1 : 0)) * kRowBytes));
        }
        return;
    }
    unsigned bytes = 0;
#pragma unroll 1
    for (int r = w; r < kTile + 1; r += W) {                  // warp-uniform
        const int q = t0 + r - 1;                             // hop-row of the signal
        float* dst = raw + r * kRawPitch;
        if (r < n_rows && row_interior(q, len, vec16)) {
            if (lane == 0) bulk_row(dst, wave_row + (long long)q * kHop, kRowBytes, bar);
            bytes += kRowBytes;
        } else {
            Sample v[3][2];
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k][0] = v[k][1] = (Sample)0;   // rows >= n_rows feed no frame that exists
            if (r < n_rows) {
                const int base = q * kHop;
#pragma unroll
                for (int k = 0; k < 3; ++k) {                     // all six loads in flight before the stores
                    const int c = lane + 32 * k;
                    if (c < 80) {
                        v[k][0] = LMFB_LDG(wave_row + reflect_index(base + 2 * c, len));
                        v[k][1] = LMFB_LDG(wave_row + reflect_index(base + 2 * c + 1, len));
                    }
                }
            }
            Sample* d = reinterpret_cast<Sample*>(dst);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int c = lane + 32 * k;
                if (c < 80) { d[2 * c] = v[k][0]; d[2 * c + 1] = v[k][1]; }
            }
        }
    }
    if (lane == 0) mbar_arrive_tx(bar, bytes);
}

// ---------------------------------------------------------------------------------------
// pass 1: the five in-register 32-point FFTs of a column, dealt round-robin to the W warps.
// The inputs come from the raw buffer: frame `lane` is rows lane and lane + 1, packed sample j at
// float 2j of that 320-sample run; input i of sub-transform n1 is the packed sample whose PFA slot
// is 32 n1 + i -- a literal offset once n1 is a literal, hence the dispatch over n1 (the loads are
// the only per-sub-transform code; the codelet and the stores are shared).
// ---------------------------------------------------------------------------------------
LMFB_CX int packed_of_slot(int s) {            // inverse of slot_of_packed
    for (int j = 0; j < kSlots; ++j) if (slot_of_packed(j) == s) return j;
    return -1;
}
LMFB_CX int raw_off_of_packed(int j) { return j < 80 ? 2 * j : kRawPitch + 2 * (j - 80); }

// (int16 samples: packed sample j is the 4-byte pair at byte 4j of the frame's first row, or of its second row)
LMFB_CX int raw_off_of_packed_i16(int j) { return j < 80 ? j : kRawPitch + (j - 80); }    // in 4-byte words

template <int N1, bool I16>
LMFB_HD void load_sub_raw(const float* __restrict__ rawl, const float2* __restrict__ win, float (&xr)[32], float (&xi)[32]) {
    static_for<0, 32>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int j = packed_of_slot(N1 * 32 + i);
        const float2 g = win[(N1 * 32 + i) * kPitch];             // same address for every lane: broadcast
        if constexpr (I16) {
            const uint32_t u = reinterpret_cast<const uint32_t*>(rawl)[raw_off_of_packed_i16(j)];
            xr[i] = (float)(int16_t)(u & 0xffffu) * g.x;
            xi[i] = (float)((int32_t)u >> 16) * g.y;
        } else {
            const float2 v = *reinterpret_cast<const float2*>(rawl + raw_off_of_packed(j));
            xr[i] = v.x * g.x; xi[i] = v.y * g.y;
        }
    });
}

//   rawl : raw + lane * kRawPitch;  col : S + lane;  win : S + kTile (the window table)
template <int W, bool I16 = false>
LMFB_HD void fft_pass1(int w, const float* __restrict__ rawl, float2* __restrict__ col, const float2* __restrict__ win) {
#pragma unroll 1
    for (int n1 = w; n1 < 5; n1 += W) {
        float xr[32], xi[32];
        switch (n1) {
            case 0:  load_sub_raw<0, I16>(rawl, win, xr, xi); break;
            case 1:  load_sub_raw<1, I16>(rawl, win, xr, xi); break;
            case 2:  load_sub_raw<2, I16>(rawl, win, xr, xi); break;
            case 3:  load_sub_raw<3, I16>(rawl, win, xr, xi); break;
            default: load_sub_raw<4, I16>(rawl, win, xr, xi); break;
        }
#ifndef LMFB_DBG_NOFFT
        fft32(xr, xi);
#endif
        float2* p = col + n1 * 32 * kPitch;
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

// adjoint of split_pair: gradients w.r.t. (xf, xp) -> gradients w.r.t. A = Z[f] and Bz = Z[160-f]
LMFB_HD void split_pair_adj(float2 gxf, float2 gxp, float sn, float cs, float2& gA, float2& gBz) {
    const float g_sr = gxf.x + gxp.x, g_si = gxf.y - gxp.y;
    const float g_Dr = gxp.x - gxf.x, g_Di = -(gxf.y + gxp.y);
    const float g_dr = fmaf(sn, g_Dr, cs * g_Di), g_di = fmaf(sn, g_Di, -cs * g_Dr);
    gA  = make_float2(g_sr + g_dr, g_si + g_di);
    gBz = make_float2(g_sr - g_dr, g_di - g_si);
}
// adjoint of a DFT (conjugate transform) = the same codelet with real and imaginary parts swapped
// on the way in and on the way out
LMFB_HD void dft5_adj(const float (&gr_)[5], const float (&gi_)[5], float (&ar)[5], float (&ai)[5]) {
    dft5(gi_, gr_, ai, ar);
}
// gradient of the masked power w.r.t. the spectrum value x (X' scale), times dP
template <int MASK>
LMFB_HD float2 masked_power_grad(float2 x, float mr, float mi, float dp) {
    if (MASK == kMaskReim) return make_float2(2.0f * mr * mr * x.x * dp, 2.0f * mi * mi * x.y * dp);
    const float m2 = MASK == kMaskPower ? 2.0f * mr * dp : 2.0f * dp;
    return make_float2(m2 * x.x, m2 * x.y);
}

// forward: masked power P' of one bin
template <int MASK>
LMFB_HD float masked_power(float2 x, float mr, float mi) {
    if (MASK == kMaskReim) { const float a = x.x * mr, b = x.y * mi; return fmaf(a, a, b * b); }
    const float p = fmaf(x.x, x.x, x.y * x.y);
    return MASK == kMaskPower ? mr * p : p;
}

// ---------------------------------------------------------------------------------------
// table fill, once per persistent CTA: every thread copies its share of the words of the kernel
// parameter and of the per-step constants (lane-indexed constant loads: a handful per thread, all
// independent).  It runs under the first tile's staging copies.  Walking the constant bank with
// warp-uniform indices and storing from lane 0 was measured at 6-8 us per CTA, a quarter of a
// one-wave launch.  (The twiddles stay correctly rounded literals: with sincospif the 1-ulp
// differences show up in the boundary-length parity test, whose imaginary parts are pure
// cancellation noise.)
// ---------------------------------------------------------------------------------------
LMFB_HD void copy_words(uint32_t* dst, const uint32_t* src, int words, int tid, int nthreads) {
    for (int i = tid; i < words; i += nthreads) dst[i] = src[i];
}
LMFB_HD void fill_steps(StepEnt* st, int tid, int nthreads) {
    for (int i = tid; i < 17 * 5; i += nthreads) {           // correctly rounded literals (generated tables)
        const int k2 = i / 5, k1 = i - 5 * k2;
        st[k2].f[k1] = kStepBin[k2][k1]; st[k2].sn[k1] = kStepSin[k2][k1]; st[k2].cs[k1] = kStepCos[k2][k1];
    }
}
// host: the same images, built once per plan for aas_lmfb_plan_upload (the kernels then fetch their
// table with ONE bulk copy instead of ~800 divergent constant-bank reads per CTA, which were measured
// at 4 us per CTA: a quarter of a one-wave launch)
inline void fill_steps_host(StepEnt* st) {
    for (int k2 = 0; k2 < 17; ++k2)
        for (int k1 = 0; k1 < 5; ++k1) {
            st[k2].f[k1] = kStepBinHost[k2][k1]; st[k2].sn[k1] = kStepSinHost[k2][k1]; st[k2].cs[k1] = kStepCosHost[k2][k1];
        }
}
inline void tables_image(FwdSmem* sm, const FwdTab& tab) {
    memset(sm, 0, sizeof(*sm));
    fill_steps_host(sm->step);
    memcpy(sm->w, tab.w, sizeof(sm->w));
    memcpy(sm->hmask, tab.hmask, sizeof(sm->hmask));
    memcpy(sm->adv, tab.adv, sizeof(sm->adv));
    memcpy(sm->row, tab.row, sizeof(sm->row));
}
inline void tables_image(BwdSmem* sm, const BwdTab& tab) {
    memset(sm, 0, sizeof(*sm));
    fill_steps_host(sm->step);
    memcpy(sm->w, tab.w, sizeof(sm->w));
    memcpy(sm->d, tab.d, sizeof(sm->d));
}

LMFB_HD void tables_fill(FwdSmem* sm, const FwdTab& tab, int tid, int nthreads) {
    fill_steps(sm->step, tid, nthreads);
    if (tab.walkable) {
        copy_words(reinterpret_cast<uint32_t*>(sm->w), reinterpret_cast<const uint32_t*>(tab.w), 2 * kBins, tid, nthreads);
        copy_words(reinterpret_cast<uint32_t*>(sm->hmask), reinterpret_cast<const uint32_t*>(tab.hmask),
                   (kGroups + 8 + kBins + 3) / 4, tid, nthreads);                // hmask and adv are adjacent in both structs
    } else {
        copy_words(sm->row, tab.row, kMaxMels, tid, nthreads);
    }
}
LMFB_HD void tables_fill(BwdSmem* sm, const BwdTab& tab, int tid, int nthreads) {
    fill_steps(sm->step, tid, nthreads);
    copy_words(reinterpret_cast<uint32_t*>(sm->w), reinterpret_cast<const uint32_t*>(tab.w), 17 * 5 * 4 + 17 * 5 * 2, tid, nthreads);
}

// ---------------------------------------------------------------------------------------
// pass 2.  Step k2 works on columns k2 and kb = (32-k2) mod 32 of the five sub-transforms: two
// 5-point DFTs -> real split -> ten bins f = (96 k1 + 65 k2) mod 160 and 160 - f -> mask.  The
// step is branch-free and identical for all k2 (the self-paired columns k2 = 0, 16, where
// kb == k2, simply compute each of their bins twice).
// ---------------------------------------------------------------------------------------
struct StepK {                                            // the per-step constants, fetched as four 128-bit words
    uint32_t f[5]; float sn[5], cs[5];
};
LMFB_HD void load_step(const StepEnt& se, StepK& k) {
#ifdef __CUDACC__
    const uint4* q = reinterpret_cast<const uint4*>(&se);
    const uint4 a = q[0], b = q[1], c = q[2], d = q[3];
    k.f[0] = a.x; k.f[1] = a.y; k.f[2] = a.z; k.f[3] = a.w; k.f[4] = b.x;
    k.sn[0] = __uint_as_float(b.y); k.sn[1] = __uint_as_float(b.z); k.sn[2] = __uint_as_float(b.w);
    k.sn[3] = __uint_as_float(c.x); k.sn[4] = __uint_as_float(c.y);
    k.cs[0] = __uint_as_float(c.z); k.cs[1] = __uint_as_float(c.w);
    k.cs[2] = __uint_as_float(d.x); k.cs[3] = __uint_as_float(d.y); k.cs[4] = __uint_as_float(d.z);
#else
    for (int i = 0; i < 5; ++i) { k.f[i] = se.f[i]; k.sn[i] = se.sn[i]; k.cs[i] = se.cs[i]; }
#endif
}
LMFB_HD void load_step_bins(const StepEnt& se, uint32_t (&f)[5]) {
#ifdef __CUDACC__
    const uint4* q = reinterpret_cast<const uint4*>(&se);
    const uint4 a = q[0];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = se.f[4];
#else
    for (int i = 0; i < 5; ++i) f[i] = se.f[i];
#endif
}

// what a step needs from global memory: index k1 -> bin f, index 5+k1 -> bin 160-f.  All pointers
// are readable for every lane (out-of-row lanes are clamped by the caller): no load is predicated.
struct StepMasks { float vr[10], vi[10]; };              // mask values (prefetched one step ahead)
struct StepD     { float d0[10], d1[10]; };              // backward: the two dE rows of each bin

template <int MASK, bool BWD, bool GW>
LMFB_HD void load_masks(const StepEnt& se, const float* __restrict__ mr, const float* __restrict__ mi, unsigned msf_bytes,
                        StepMasks& in) {
    uint32_t f[5];
    load_step_bins(se, f);
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const uint32_t fp = kBins - 1 - f[k1];
        if (LMFB_NEEDS_MASK_R(MASK, BWD, GW)) { in.vr[k1] = LMFB_LDG(at_row(mr, f[k1], msf_bytes)); in.vr[5 + k1] = LMFB_LDG(at_row(mr, fp, msf_bytes)); }
        if (LMFB_NEEDS_MASK_I(MASK, BWD, GW)) { in.vi[k1] = LMFB_LDG(at_row(mi, f[k1], msf_bytes)); in.vi[5 + k1] = LMFB_LDG(at_row(mi, fp, msf_bytes)); }
    }
}

// issued at the start of the step that consumes it: the two 5-point DFTs and the split (about
// 200 instructions that need nothing from global memory) run while these loads are in flight.
// de1 = dE + one row.
LMFB_HD void load_d(const uint32_t (*drow)[2], const float* __restrict__ dE, const float* __restrict__ de1,
                    unsigned sem_bytes, StepD& in, const float* __restrict__ de_s = nullptr) {
    if (de_s) {                                              // the tile's dE rows are in shared memory (this lane's column)
#pragma unroll
        for (int k1 = 0; k1 < 5; ++k1) {
            const float* pf = de_s + drow[k1][0] * 32u;
            const float* pp = de_s + drow[k1][1] * 32u;
            in.d0[k1] = pf[0]; in.d1[k1] = pf[32];
            in.d0[5 + k1] = pp[0]; in.d1[5 + k1] = pp[32];
        }
        return;
    }
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const uint32_t df = drow[k1][0], dp = drow[k1][1];
        in.d0[k1]     = LMFB_LDG(at_row(dE, df, sem_bytes));
        in.d1[k1]     = LMFB_LDG(at_row(de1, df, sem_bytes));
        in.d0[5 + k1] = LMFB_LDG(at_row(dE, dp, sem_bytes));
        in.d1[5 + k1] = LMFB_LDG(at_row(de1, dp, sem_bytes));
    }
}

//   col : this lane's float2 column (S + lane);  pl : this lane's float column ((float*)S + lane)
//   forward : masked power of bin f -> pl[p_off(f)], i.e. slot row f, one of the rows this step
//             has just read and no other step touches
//   backward: gradients = (2 Mr Re'^2, 2 Mi Im'^2) * dP, or (Re'^2 + Im'^2) * dP for 'power',
//             stored to row f of gr/gi; every lane's pointers are valid, `inrow` gates the store
template <int MASK, bool BWD, bool GW, class SM>
LMFB_HD void pass2_step(int k2, float2* __restrict__ col, float* __restrict__ pl, const SM& sm, const StepMasks& in,
                        const float* __restrict__ dE, const float* __restrict__ de1, unsigned sem_bytes, unsigned msf_bytes,
                        float* __restrict__ gr, float* __restrict__ gi, bool inrow, const float* __restrict__ de_s = nullptr) {
    StepD d;
    if constexpr (BWD && MASK != kStftOut) load_d(sm.d[k2], dE, de1, sem_bytes, d, de_s);
    const int kb = (32 - k2) & 31;
    const float2* ca = col + k2 * kPitch;
    const float2* cb = col + kb * kPitch;
    float ar[5], ai[5], br[5], bi[5], Ar[5], Ai[5], Br[5], Bi[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) {
        const float2 v = ca[n * 32 * kPitch]; ar[n] = v.x; ai[n] = v.y;
        const float2 u = cb[n * 32 * kPitch]; br[n] = u.x; bi[n] = u.y;
    }
    // forward: the compact power rows written below overlap the float2 words OTHER lanes of this warp
    // have just read (same slot rows); order the warp's reads before its writes
    if constexpr (!BWD) LMFB_SYNCWARP();
    StepK k;
    load_step(sm.step[k2], k);
#ifndef LMFB_DBG_NOFFT
    dft5(ar, ai, Ar, Ai);
    dft5(br, bi, Br, Bi);
#else
#pragma unroll
    for (int n = 0; n < 5; ++n) { Ar[n] = ar[n]; Ai[n] = ai[n]; Br[n] = br[n]; Bi[n] = bi[n]; }
#endif
    // GW (gradient into the waveform): the gradients w.r.t. the two columns' DFT5 outputs
    float gAr[5], gAi[5], gBr[5], gBi[5];
    const bool self = kb == k2;                              // columns 0 and 16 pair with themselves
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
        const int kp = (5 - k1) % 5;
        const uint32_t f = k.f[k1], fp = kBins - 1 - f;
        float2 xf, xp;
        split_pair(Ar[k1], Ai[k1], Br[kp], Bi[kp], k.sn[k1], k.cs[k1], xf, xp);
        if constexpr (!BWD) {
            pl[f * kRow] = masked_power<MASK>(xf, in.vr[k1], in.vi[k1]);
            pl[fp * kRow] = masked_power<MASK>(xp, in.vr[5 + k1], in.vi[5 + k1]);    // (bin 160: the row behind the slots)
        } else if constexpr (MASK == kStftOut) {
            // the spectrum is carried as X' = 2X; `dE` doubles as nothing here, the scale (0.5 for
            // frames that exist, 0 beyond) arrives in the first mask slot
            const float sc = in.vr[0];
            st_if(at_row(gr, f, msf_bytes), sc * xf.x, inrow);
            st_if(at_row(gi, f, msf_bytes), sc * xf.y, inrow);
            st_if(at_row(gr, fp, msf_bytes), sc * xp.x, inrow);
            st_if(at_row(gi, fp, msf_bytes), sc * xp.y, inrow);
        } else {
#ifdef __CUDACC__
            const float4 wv = *reinterpret_cast<const float4*>(sm.w[k2][k1]);
            const float wlf = wv.x, whf = wv.y, wlp = wv.z, whp = wv.w;
#else
            const float wlf = sm.w[k2][k1][0], whf = sm.w[k2][k1][1], wlp = sm.w[k2][k1][2], whp = sm.w[k2][k1][3];
#endif
            const float dpf = fmaf(whf, d.d1[k1], wlf * d.d0[k1]);
            const float dpp = fmaf(whp, d.d1[5 + k1], wlp * d.d0[5 + k1]);
            if constexpr (GW) {
                // in a self-paired column every bin shows up twice (outputs 3, 4 repeat 2, 1) and
                // bin 80 is its own partner: count each bin once
                const bool dup = self && k1 >= 3, nopartner = f == fp;
                // (lanes past the end of the row read a clamped dE column: they are frames that do not exist)
                float2 gxf = masked_power_grad<MASK>(xf, in.vr[k1], in.vi[k1], (dup || !inrow) ? 0.0f : dpf);
                float2 gxp = masked_power_grad<MASK>(xp, in.vr[5 + k1], in.vi[5 + k1], (dup || nopartner || !inrow) ? 0.0f : dpp);
                float2 gA, gBz;
                split_pair_adj(gxf, gxp, k.sn[k1], k.cs[k1], gA, gBz);
                gAr[k1] = gA.x; gAi[k1] = gA.y; gBr[kp] = gBz.x; gBi[kp] = gBz.y;
            }
            if (MASK == kMaskNone) {
                // nothing to store: only the waveform takes a gradient
            } else if (MASK == kMaskReim) {
                st_if_cs(at_row(gr, f, msf_bytes), 2.0f * in.vr[k1] * xf.x * xf.x * dpf, inrow);
                st_if_cs(at_row(gi, f, msf_bytes), 2.0f * in.vi[k1] * xf.y * xf.y * dpf, inrow);
                st_if_cs(at_row(gr, fp, msf_bytes), 2.0f * in.vr[5 + k1] * xp.x * xp.x * dpp, inrow);
                st_if_cs(at_row(gi, fp, msf_bytes), 2.0f * in.vi[5 + k1] * xp.y * xp.y * dpp, inrow);
            } else {
                st_if_cs(at_row(gr, f, msf_bytes), fmaf(xf.x, xf.x, xf.y * xf.y) * dpf, inrow);
                st_if_cs(at_row(gr, fp, msf_bytes), fmaf(xp.x, xp.x, xp.y * xp.y) * dpp, inrow);
            }
        }
    }
    if constexpr (BWD && GW) {
        // adjoint 5-point DFTs, in place: the slots of these two columns now hold the gradient w.r.t.
        // the pass-1 outputs (self-paired column: both halves belong to the same column)
#pragma unroll
        for (int n = 0; n < 5; ++n) if (self) { gAr[n] += gBr[n]; gAi[n] += gBi[n]; }
        float ar2[5], ai2[5];
        dft5_adj(gAr, gAi, ar2, ai2);
        float2* wa = col + k2 * kPitch;
#pragma unroll
        for (int n = 0; n < 5; ++n) wa[n * 32 * kPitch] = make_float2(ar2[n], ai2[n]);
        if (!self) {
            dft5_adj(gBr, gBi, ar2, ai2);
            float2* wb = col + kb * kPitch;
#pragma unroll
            for (int n = 0; n < 5; ++n) wb[n * 32 * kPitch] = make_float2(ar2[n], ai2[n]);
        }
    }
}

// GW, after pass 2 and a block barrier: adjoint of pass 1 -- conjugate 32-point transforms, then
// the window -- leaves in every column the gradient w.r.t. that frame's 160 packed raw samples
template <int W>
LMFB_HD void fft_pass1_adj(int w, float2* __restrict__ col, const float2* __restrict__ win) {
#pragma unroll 1
    for (int n1 = w; n1 < 5; n1 += W) {
        float2* p = col + n1 * 32 * kPitch;
        const float2* wn = win + n1 * 32 * kPitch;
        float xr[32], xi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float2 v = p[i * kPitch]; xr[i] = v.y; xi[i] = v.x; }   // swapped in
        fft32(xr, xi);
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float2 g = wn[i * kPitch]; p[i * kPitch] = make_float2(xi[i] * g.x, xr[i] * g.y); }   // swapped out
    }
}

#ifdef __CUDACC__
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add2(float* p, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}
#else
static inline void red_add(float* p, float v) { *p += v; }
static inline void red_add2(float* p, float2 v) { p[0] += v.x; p[1] += v.y; }
#endif

// GW, after another barrier: adjoint of the staging (overlap-add).  Hop-row r collects the first
// half of frame r and the second half of frame r-1 and is ADDED to the (zero-initialised) waveform
// gradient: neighbouring tiles share their edge rows, and the reflect padding folds the two ends of
// an utterance back onto its first and last samples.
template <int W>
LMFB_HD void unstage_tile(int w, int lane, const StageLane& sl, float* __restrict__ gwave_row, int len,
                          int t0, int n_rows, const float2* __restrict__ S, bool vec_ok) {
    constexpr int kShare = (kTile + 1 + W - 1) / W;
    const int r_lo = w * kShare;
    const int r_hi = r_lo + kShare < kTile + 1 ? r_lo + kShare : kTile + 1;
#pragma unroll 1
    for (int r = r_lo; r < r_hi && r < n_rows; ++r) {
        const int q = t0 + r - 1;
        const bool fast = row_interior(q, len, vec_ok);
        // a row whose two frames both belong to this tile, and which the reflect padding of neither end of the
        // utterance folds back onto, has no other writer: a plain store (31 of a tile's 33 rows; the atomics
        // were 5,280 L2 read-modify-writes per tile)
        const bool solo = fast && r >= 1 && r < kTile && q >= 2 && (q + 3) * kHop <= len;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int c = lane + 32 * k;
            if (c < 80) {
                float2 v = make_float2(0.0f, 0.0f);
                if (r < kTile) { const float2 a = S[sl.slot_a[k] + r]; v.x += a.x; v.y += a.y; }
                if (r >= 1)    { const float2 b = S[sl.slot_b[k] + r - 1]; v.x += b.x; v.y += b.y; }
                if (solo) {
                    *reinterpret_cast<float2*>(gwave_row + (long long)q * kHop + 2 * c) = v;
                } else if (fast) {
                    red_add2(gwave_row + (long long)q * kHop + 2 * c, v);
                } else {
                    red_add(gwave_row + reflect_index(q * kHop + 2 * c, len), v.x);
                    red_add(gwave_row + reflect_index(q * kHop + 2 * c + 1, len), v.y);
                }
            }
        }
    }
}

// pass 2 over the 17 column pairs, dealt round-robin to the W warps.  The global inputs of a warp's
// steps are loaded AHEAD register sets in advance (AHEAD = 1: ping-pong; AHEAD = 2: three sets in
// rotation, the loop body is unrolled over the rotation so that no register ever moves): bytes in
// flight per warp are what bounds these kernels (throughput = bytes in flight / ~2000 cycles of
// loaded memory latency), and these are the only registers that can buy any.
// `m[0 .. AHEAD-1]` must already hold the masks of the warp's first AHEAD steps (k2 = w, w + W, ...):
// the caller issues those loads before the block barrier that ends pass 1.
template <int AHEAD>
struct MaskSets { StepMasks m[AHEAD + 1]; };
#ifndef LMFB_AHEAD_FWD
#  define LMFB_AHEAD_FWD 1
#endif
#ifndef LMFB_AHEAD_BWD
#  define LMFB_AHEAD_BWD 1
#endif
constexpr int kAheadFwd = LMFB_AHEAD_FWD, kAheadBwd = LMFB_AHEAD_BWD;

template <int W, int MASK, bool BWD, int AHEAD, bool GW, class SM>
LMFB_HD void preload_masks(int w, const SM& sm, const float* __restrict__ mr, const float* __restrict__ mi,
                           unsigned msf_bytes, MaskSets<AHEAD>& ms) {
#pragma unroll
    for (int i = 0; i < AHEAD; ++i)
        if (w + i * W <= 16) load_masks<MASK, BWD, GW>(sm.step[w + i * W], mr, mi, msf_bytes, ms.m[i]);
}

template <int W, int MASK, bool BWD, int AHEAD, bool GW, class SM>
LMFB_HD void fft_pass2(int w, float2* __restrict__ col, float* __restrict__ pl, const SM& sm, MaskSets<AHEAD>& ms,
                       const float* __restrict__ mr, const float* __restrict__ mi,
                       const float* __restrict__ dE, unsigned sem_bytes, unsigned msf_bytes,
                       float* __restrict__ gr, float* __restrict__ gi, bool inrow, const float* __restrict__ de_s = nullptr) {
    const float* de1 = at_row(dE, 1u, sem_bytes);
    constexpr int R = AHEAD + 1;                            // register sets in rotation
#pragma unroll 1
    for (int k2 = w; k2 <= 16; k2 += R * W) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int k = k2 + i * W;                       // this step; its masks are in set i
            if (k <= 16) {
                const int kn = k + AHEAD * W;               // the step AHEAD later goes to set (i + AHEAD) mod R
                if (kn <= 16) load_masks<MASK, BWD, GW>(sm.step[kn], mr, mi, msf_bytes, ms.m[(i + AHEAD) % R]);
                pass2_step<MASK, BWD, GW>(k, col, pl, sm, ms.m[i], dE, de1, sem_bytes, msf_bytes, gr, gi, inrow, de_s);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// phase 3 (forward).  Every warp OWNS a run of whole filters (FwdTab::lo .. hi, a cost-balanced
// partition made on the host) and walks, in ascending order with two running sums (filters ml(f),
// ml(f) + 1), every bin that feeds one of them; when the band moves on, the finished sum goes to a
// row of the warp's own.  Neighbouring warps both walk the bins between their two border filters
// (one inter-centre interval: ~10 % more bins in total), which buys: no partial sums, no block barrier
// between the walk and the log1p -- a lane reads back only what it wrote itself -- and one
// shared-memory load per filter in the finish.  (History: bins dealt in equal shares, partial sums
// in W row sets, a barrier, and a finish that added up to W partial rows per filter: 4.2 k warp
// instructions per tile and 8 % of the kernel's stall samples on that barrier.)
// Rows: filter m of warp w is slot row 1 + 2w + m (second half): the row sets of neighbouring warps,
// which overlap by the two border filters, never collide.
//   out : out + n*stride_n + t (row m at + m*som); inrow: t < Tmax; valid: t < T_i
// ---------------------------------------------------------------------------------------
struct Walk { float acc0, acc1; float* cur; };

// one bin; `h` (warp-uniform): the band moves on by one filter before this bin.  A real branch: three
// bins in four take the short path (two FMAs); the predicated, branch-free form (store, two selects and a
// pointer select per bin) was 1 - 2 % slower in the forward.
LMFB_HD void walk_bin(Walk& wk, float p, float2 wgt, bool h) {
    if (h) {
        sts_if_noalias(wk.cur, wk.acc0, true);
        wk.acc0 = wk.acc1; wk.acc1 = 0.0f; wk.cur += kRow;
    }
    wk.acc0 = fmaf(wgt.x, p, wk.acc0);
    wk.acc1 = fmaf(wgt.y, p, wk.acc1);
}
// general hand-over of `adv` filters (bases in which several filters end on the same bin)
LMFB_HD void walk_bin_multi(Walk& wk, float p, float2 wgt, unsigned adv) {
    if (adv != 0) {
        sts_if_noalias(wk.cur, wk.acc0, true); wk.cur += kRow;
        wk.acc0 = wk.acc1; wk.acc1 = 0.0f;
        if (adv > 1) {
            sts_if_noalias(wk.cur, wk.acc0, true); wk.cur += kRow;
            wk.acc0 = 0.0f;
#pragma unroll 1
            for (unsigned i = 2; i < adv; ++i) { sts_if_noalias(wk.cur, 0.0f, true); wk.cur += kRow; }
        }
    }
    wk.acc0 = fmaf(wgt.x, p, wk.acc0);
    wk.acc1 = fmaf(wgt.y, p, wk.acc1);
}

// hand-over bits of the eight bins f .. f + 7 (bit i: the band moves on before bin f + i)
LMFB_HD unsigned hand_bits(const uint8_t* hmask, int f) {
    const uint32_t* hb = reinterpret_cast<const uint32_t*>(hmask);
    const unsigned long long two = ((unsigned long long)hb[(f >> 5) + 1] << 32) | hb[f >> 5];
    return (unsigned)(two >> (f & 31)) & 0xffu;
}

// the walk over this warp's bins; leaves a sum in row er + m * kRow for every filter m0 .. m1
LMFB_HD void phase3_walk(int w, const float* __restrict__ pl, float* __restrict__ er, const FwdSmem& sm, const FwdTab& tab) {
    const int b0 = tab.b0[w], b1 = tab.b1[w];
    if (b0 >= b1) return;
    Walk wk;
    wk.acc0 = wk.acc1 = 0.0f;
    wk.cur = er + (int)tab.m0[w] * kRow;
    if (!tab.multi) {
        // eight bins at a time: all sixteen loads of a group are in flight before the hand-over
        // chain starts.  (A piece-by-piece walk with counted loops and no per-bin predicates
        // was measured: 20 % fewer instructions in this phase, but twice its duration -- short dependent
        // load -> FMA loops leave a warp nothing to overlap; what a phase costs is its latency.)
        const float* pp = pl + b0 * kRow;
        const float2* wp = sm.w + b0;
        int f = b0;
#pragma unroll 1
        for (; f + 8 <= b1; f += 8) {
            float p[8]; float2 wg[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = pp[i * kRow]; wg[i] = wp[i]; }
            unsigned hm = hand_bits(sm.hmask, f);
            if (f == b0) hm &= ~1u;                        // nothing is complete before the warp's first bin
#pragma unroll
            for (int i = 0; i < 8; ++i) walk_bin(wk, p[i], wg[i], (hm >> i) & 1u);
            pp += 8 * kRow; wp += 8;
        }
        if (f < b1) {                                      // the last 1 .. 7 bins
            const int rem = b1 - f;
            float p[7]; float2 wg[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                p[i] = 0.0f; wg[i] = make_float2(0.0f, 0.0f);
                if (i < rem) { p[i] = pp[i * kRow]; wg[i] = wp[i]; }
            }
            unsigned hm = hand_bits(sm.hmask, f);
            if (f == b0) hm &= ~1u;
#pragma unroll
            for (int i = 0; i < 7; ++i) if (i < rem) walk_bin(wk, p[i], wg[i], (hm >> i) & 1u);
        }
    } else {
#pragma unroll 1
        for (int f = b0; f < b1; ++f) walk_bin_multi(wk, pl[f * kRow], sm.w[f], f == b0 ? 0u : sm.adv[f]);
    }
    sts_if_noalias(wk.cur, wk.acc0, true);
    sts_if_noalias(wk.cur + kRow, wk.acc1, true);
}

// the sum of filter m as the walk of this warp left it (filters no bin of the walk reached are empty)
LMFB_HD float walked_sum(const float* __restrict__ er, int m, int m0, int m1) {
    return (m >= m0 && m <= m1) ? er[m * kRow] : 0.0f;
}

template <int W>
LMFB_HD void phase3_own(int w, float* __restrict__ pl, const FwdSmem& sm, const FwdTab& tab,
                        float* __restrict__ out, unsigned som_bytes, bool inrow, bool valid,
                        bool first = true, bool last = true) {
    float* er = pl + e_off(2 * w);
    phase3_walk(w, pl, er, sm, tab);
#ifdef __CUDACC__
    asm volatile("" ::: "memory");                         // the rows are written behind the optimiser's back (sts_if_noalias)
#endif
    const int m_lo = tab.lo[w], m_end = (int)tab.hi[w] + 1;
    const int m0 = tab.m0[w], m1 = tab.m1[w];
    if (m_lo >= m_end) return;
    float* op = at_row(out, (uint32_t)m_lo, som_bytes);
    if (first && last) {
        // single channel (the common case): four filters in flight -- the log1p chains are what this
        // part waits for.  The last group is padded with filters that are not stored rather than handed
        // to a one-at-a-time loop (same instructions, a quarter of the latency).
#pragma unroll 1
        for (int m = m_lo; m < m_end; m += 4) {
            float e[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = walked_sum(er, m + i, m0, m + i < m_end ? m1 : -1);
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = valid ? log1pf(e[i]) : 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { st_if(op, e[i], inrow && m + i < m_end); op = at_row(op, 1u, som_bytes); }
        }
        return;
    }
#pragma unroll 1
    for (int m = m_lo; m < m_end; ++m) {
        float e = walked_sum(er, m, m0, m1);
        if (!first) e += inrow ? *op : 0.0f;                 // multi-channel: partial sums of E travel through `out`
        const float y = last ? (valid ? log1pf(e) : 0.0f) : e;
        st_if(op, y, inrow);
        op = at_row(op, 1u, som_bytes);
    }
}

// ---------------------------------------------------------------------------------------
// phase 3 for ANY other basis (the reference applies whatever (M, F) matrix it is given with a k=1
// conv1d, model.py:167, :196): a gather.  Filters are dealt round-robin to the W warps; a warp
// forms E[m] = 1/4 sum_f B[m, f] P'[f] over the filter's row [lo, lo + cnt) with the weights read
// from the caller's device copy of the matrix (same address in every lane: a broadcast load),
// applies log1p and stores the row segment.  One pass, no partial sums.  Measured on the default
// basis (weights in shared memory, two bins per load): 7 % slower than walk + finish -- the per-filter
// chains are latency-bound -- which is why triangular filterbanks keep the walk.
//   mel : device (M, 161) matrix
//   first / last : multi-channel input: channel 0 / the last channel; in between the partial sums
//         of E travel through `out` (the same thread writes and reads them)
// ---------------------------------------------------------------------------------------
template <int W>
LMFB_HD void phase3_gather(int w, const float* __restrict__ pl, const FwdSmem& sm, int n_mels,
                           const float* __restrict__ mel, float* __restrict__ out, unsigned som_bytes,
                           bool inrow, bool valid, bool first, bool last) {
    float* op = at_row(out, (uint32_t)w, som_bytes);
#pragma unroll 1
    for (int m = w; m < n_mels; m += W) {
        const uint32_t d = sm.row[m];
        const int lo = (int)(d & 255u), c = (int)(d >> 8);
        const float* pp = pl + lo * kRow;
        const float* wr = mel + m * kBins + lo;
        float a0 = 0.0f, a1 = 0.0f;
        int i = 0;
#pragma unroll 1
        for (; i + 2 <= c; i += 2) {
            a0 = fmaf(LMFB_LDG(wr + i), pp[0], a0);
            a1 = fmaf(LMFB_LDG(wr + i + 1), pp[kRow], a1);
            pp += 2 * kRow;
        }
        if (i < c) a0 = fmaf(LMFB_LDG(wr + i), pp[0], a0);
        float e = 0.25f * (a0 + a1);
        if (!first) e += inrow ? *op : 0.0f;
        const float y = last ? (valid ? log1pf(e) : 0.0f) : e;
        st_if(op, y, inrow);
        op = at_row(op, (uint32_t)W, som_bytes);
    }
}

// L2 prefetch of the row segments a tile will read: threads take rows idx, idx+cnt, ...; a
// 128-byte segment may straddle two lines, so both ends are touched.  Issued when the tile starts.
// Measured alternatives on 256 x 10 s (forward / backward ms): this 0.194 / 0.204; none 0.212 / 0.242;
// a tile AHEAD (when the next tile's rows are requested) 0.201 / 0.246 -- the lines are gone again by the
// time they are used (the L2 turns over in ~12 us under this traffic); step by step, two steps ahead of
// the loads, 0.199 / 0.213 (the extra live values spill).  (Touching all five sectors
// of a row was measured: slower, the extra prefetches cost more load/store-unit time than they save.)
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

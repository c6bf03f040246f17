// B200 (sm_100a) kernels and C ABI of the LMFB front-end.  See include/aas_lmfb.h for the
// contract and lmfb_core.cuh for the per-thread algorithm.
//
//   K1  lmfb_k1<MASK, BWD, W, CTAS>   W warps per tile of 32 frames (lane = frame): stage wave ->
//                            320-pt real FFT -> mask -> banded mel -> log1p      (forward)
//                            or -> d mask from dE                                 (backward)
//                            or -> the spectrum itself, (N, 2*161, T)             (aas_lmfb_stft);
//                            tiles handed out by cluster launch control.
//   K2  cmvn_fwd* / cmvn_bwd*  per-utterance mean/variance normalisation of the (M, T) rows
//                            and its gradient folded with d log1p.
//   K3  l1_abs_*             L1Loss_mask: deterministic sum |a - b| and its gradient.
//
// No cuFFT, no library calls, no CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <mutex>
#include <set>
#include <utility>

#include "../../include/aas_lmfb.h"
#include "lmfb_core.cuh"
#include "mel_band.hpp"

namespace aas_lmfb {

// division of a non-negative int < 2^31 by a fixed positive divisor as one multiply-high and a shift
// (the tile decode runs in every warp of every tile: two signed divisions were ~60 instructions)
struct FastDiv {
    unsigned m; int s;                      // s < 0: the divisor is 1
    __host__ __device__ int div(int n) const {
#ifdef __CUDA_ARCH__
        return s < 0 ? n : (int)(__umulhi((unsigned)n, m) >> s);
#else
        return s < 0 ? n : (int)(((unsigned long long)(unsigned)n * m) >> (32 + s));
#endif
    }
};
static inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    if (d <= 1) { f.m = 0; f.s = -1; return f; }
    int c = 0;                              // c = ceil(log2 d)
    while ((1LL << c) < d) ++c;
    f.s = c - 1;
    f.m = (unsigned)((((unsigned long long)1 << (32 + f.s)) + (unsigned)d - 1) / (unsigned)d);
    return f;
}

struct K1Args {
    const void*    wave;       // fp32 samples, or int16 PCM in the I16 kernels
    const int32_t* lengths;
    long long      wave_stride;
    const float*   mask_r;
    const float*   mask_i;
    long long      msn;
    unsigned       msf;
    const float*   window;
    float*         out;        // forward: (N, M, Tmax), receives log1p(E)
    const float*   dE;         // backward: (N, M, Tmax)
    float*         gr;
    float*         gi;
    int            tmax;
    int            tiles_per_utt;
    int            total_tiles;
    int            vec_ok;
    long long*     timeline;   // debug builds (-DLMFB_TIMELINE) only: per-warp phase clocks
    int            dynamic;    // tiles handed out by cluster launch control (grid = one CTA per tile)
    float*         gwave;      // backward with GW: (N, wave_stride) gradient w.r.t. the samples, zeroed by the caller of the kernel
    const float*   mel;        // forward, generic basis only: device copy of the (M, 161) matrix
    int            n_ch;       // channels per utterance (model.py:167: the basis repeats over channels, i.e. power sums)
    long long      wave_stride_ch;   // samples between the channels of an utterance
    long long      wave_len;   // samples of a row that may be read (lengths are clamped to it); 0: trust lengths
    const void*    tab_dev;    // device image of the shared-memory table (aas_lmfb_plan_upload), or NULL: fill from the parameter
    FastDiv        div_tpu, div_nch;   // / tiles_per_utt, / n_ch
    int32_t*       frame_lens; // forward, optional: (N,) receives T_n = 1 + lengths[n] / 160 clipped to Tmax (0 for an empty utterance)
};


// ---- dynamic tile scheduling with cluster launch control (sm_100) -------------------------------
// The grid has one CTA per tile, but only the resident CTAs ever run: a running CTA asks the
// hardware to CANCEL a CTA that has not been launched yet and, when that succeeds, does that CTA's
// tile itself (persistent CTAs without a global work counter: nothing to allocate or reset, safe
// across streams).  Measured with a static round-robin deal of the tiles to 740 persistent CTAs:
// the CTAs of one launch finish between 163 and 212 us (mean 185): 13 % of the kernel is spent
// waiting for the slowest SMs.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LMFB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LMFB_DONE;\n\t"
        "bra LMFB_WAIT;\n\t"
        "LMFB_DONE:\n\t"
        "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one thread: request the cancellation of a pending CTA; 16 bytes of answer land in *resp and complete *bar
__device__ __forceinline__ void clc_request(uint4* resp, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" :: "r"(smem_u32(bar)) : "memory");
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                 :: "r"(smem_u32(resp)), "r"(smem_u32(bar)) : "memory");
}
// every thread: was a CTA cancelled for us, and which one
__device__ __forceinline__ bool clc_answer(const uint4* resp, int& ctaid_x) {
    unsigned ok, x;
    asm volatile(
        "{\n\t"
        ".reg .pred p1;\n\t"
        ".reg .b128 r;\n\t"
        "mov.u32 %1, 0;\n\t"
        "ld.shared.b128 r, [%2];\n\t"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
        "selp.u32 %0, 1, 0, p1;\n\t"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, r;\n\t"
        "}" : "=r"(ok), "=r"(x) : "r"(smem_u32(resp)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // our read before the next asynchronous write
    ctaid_x = (int)x;
    return ok != 0;
}

// where a (tile, channel) unit lives: decoded once, used by the staging request and by the passes
struct Unit {
    int tile, ch;          // tile index of the launch, channel inside the utterance
    int n, t0, len, T;     // utterance, first frame, samples, frames
    bool real;             // false: the tile lies entirely in the zero padding (nothing to stage)
};

template <bool BWD>
__device__ __forceinline__ void decode_unit(const K1Args& a, int tile, int ch_fwd, Unit& u) {
    // forward: a tile belongs to an utterance and loops over its channels (the power sums);
    // backward: every (utterance, channel) has tiles of its own (they share only dE)
    const int nn = a.div_tpu.div(tile);
    u.tile = tile;
    u.t0 = (tile - nn * a.tiles_per_utt) * kTile;
    u.n  = BWD ? a.div_nch.div(nn) : nn;
    u.ch = BWD ? nn - u.n * a.n_ch : ch_fwd;
    int len = a.lengths[u.n];
    if (a.wave_len > 0 && (long long)len > a.wave_len) len = (int)a.wave_len;
    u.len = len;
    int T = len >= 1 ? 1 + len / kHop : 0;
    u.T = T < a.tmax ? T : a.tmax;
    u.real = u.t0 < u.T;
}

// (Which warps take the 17 mod W longer shares of pass 2 does not matter: giving them to the highest warp
// indices instead of the lowest was measured, 0.4385 vs 0.4345 ms per step on 256 x 10 s.)
// Measured and rejected in this kernel (256 x 10 s step, ms; builds with and without each change compared back
// to back on the same box, each twice):
//   * half steps for the self-paired columns 0 and 16 of pass 2 (one 5-point DFT, three outputs; the 17 steps
//     then weigh 3 | 3 | 3 | 3.55 | 3.55 on five warps and 2 .. 2.1 on eight), tested inside the step loop and
//     peeled in front of it: 0.4465 vs 0.4345, and 24.8 vs 22.8 us for the eight-warp forward on 30 x 6 s.
//     Less work, and slower, in every build tried;
//   * the next tile's FFT in registers BEFORE the barrier that frees the scratch (warps that finish a tile early
//     start their next FFT instead of waiting; needs two answer slots for the launch-control unit and the dE
//     staging behind that barrier): forward unchanged (0.1848 vs 0.1845), backward 0.2094 vs 0.1951.  With
//     three thread blocks per SM the stall samples on a block barrier are not idle issue slots;
//   * step constants and mel weights read from constant memory with the (now warp-uniform) step index instead
//     of from the shared-memory image: forward 0.1950 vs 0.1846, backward 0.1992 vs 0.1931;
//   * tile index and utterance length made warp-uniform with a shuffle (like the warp index): no change;
//   * a 16-instruction log1p for non-negative arguments (2 atanh(x / (2 + x)) below 1/2, hardware log2 above)
//     in place of log1pf (31 instructions, a tenth of the forward kernel's): 0.4135 - 0.4183 vs 0.4158 - 0.4163;
//     the phase is bound by its dependent chains, not by issue slots.
template <int MASK, bool BWD, int W, int CTAS, bool GW = false, bool I16 = false>
__global__ void __launch_bounds__(kTile * W, CTAS)
lmfb_k1(const __grid_constant__ K1Args a, const __grid_constant__ typename TabOf<BWD>::Param tab) {
    typedef typename TabOf<BWD>::Smem SM;
    // mask register sets loaded ahead in pass 2.  One for both directions: two (three sets in
    // rotation) were measured, forward 198 vs 193 us, backward 288 vs 276 us on 256 x 10 s.
    constexpr int AHEAD = BWD ? kAheadBwd : kAheadFwd;
    extern __shared__ __align__(16) float2 S[];
    SM& sm = *reinterpret_cast<SM*>(reinterpret_cast<char*>(S) + kScratchBytes);
    // behind the tables: the answer of the launch-control unit, its barrier, the raw buffer's barrier,
    // and the raw buffer itself
    uint4*    clc_resp = reinterpret_cast<uint4*>(reinterpret_cast<char*>(&sm) + TabBytes<BWD>::value);
    uint64_t* clc_bar  = reinterpret_cast<uint64_t*>(clc_resp + 1);
    uint64_t* raw_bar  = clc_bar + 1;
    uint64_t* tab_bar  = clc_bar + 2;
    float*    raw      = reinterpret_cast<float*>(reinterpret_cast<char*>(clc_resp) + kCtlBytes);
    float*    de_buf   = raw + kRawBytes / 4;                  // backward: the tile's dE rows
    const bool de_smem = BWD && MASK != kStftOut && tab.n_mels <= kDeSmemRows;
#ifdef LMFB_TIMELINE
    unsigned long long gt_entry;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_entry));
#endif
    const int lane = threadIdx.x & 31;
    // (through a shuffle from lane 0: the optimiser then knows that the warp index is the same in all lanes and
    // keeps what is derived from it -- the operands of the bulk copies, table rows -- in uniform registers)
    const int w    = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    // The range of the warp index decides how much of the per-warp dealing (rows of the staging, sub-transforms,
    // pass-2 steps) folds at compile time.  The optimiser derives it from __launch_bounds__ in some builds and
    // not in others (same source: parallel vs single-module compilation of this file gave 4,800 vs 4,896 vs
    // 5,120 instructions for the default forward kernel, 0.189 vs 0.195 ms on 256 x 10 s): say it.
    __builtin_assume(threadIdx.x < (unsigned)(kTile * W));
    __builtin_assume(w < W);
    const int n_mels = tab.n_mels;
    float2* col = S + lane;
    float*  pl  = reinterpret_cast<float*>(S) + lane;
    const float* rawl = raw + lane * kRawPitch;
    const unsigned msf_bytes = a.msf * 4u;
    const unsigned som = (unsigned)a.tmax, som_bytes = som * 4u;
    const int nch_fwd = BWD ? 1 : a.n_ch;                      // channel units per forward tile
    unsigned clc_phase = 0, raw_phase = 0;

    if (threadIdx.x == 0) {
        mbar_init(clc_bar, 1);
        mbar_init(raw_bar, W);                                 // one arrival per warp and staged unit
        mbar_init(tab_bar, 1);
        if (a.tab_dev) {                                       // the table image: one bulk copy
            mbar_arrive_tx(tab_bar, (unsigned)TabBytes<BWD>::value);
            bulk_copy(&sm, a.tab_dev, (unsigned)TabBytes<BWD>::value, tab_bar);
        }
    }
    __syncthreads();

    // request the rows of a unit (no-op for a padding tile)
    auto stage = [&](const Unit& u) {
        if (!u.real) return;
        const long long woff = (long long)u.n * a.wave_stride + (long long)u.ch * a.wave_stride_ch;
        const void* wave_row = static_cast<const char*>(a.wave) + woff * (I16 ? 2 : 4);
        const int n_rows = (u.T - u.t0 < kTile ? u.T - u.t0 : kTile) + 1;      // hop-rows that feed a valid frame
        stage_raw<W, I16>(w, lane, wave_row, u.len, u.t0, n_rows, raw, raw_bar, (a.vec_ok & 1) != 0);
    };

    // L2 prefetch of the mask rows (and dE rows) a unit will read in pass 2
    auto prefetch_unit = [&](const Unit& u) {
        if (!u.real) return;
        const long long mrow = (long long)u.n * a.msn + (long long)u.ch * kBins * a.msf;
        if (LMFB_NEEDS_MASK_R(MASK, BWD, GW)) prefetch_rows_l2(threadIdx.x, kTile * W, a.mask_r + mrow, a.msf, kBins, u.t0, a.tmax);
        if (LMFB_NEEDS_MASK_I(MASK, BWD, GW)) prefetch_rows_l2(threadIdx.x, kTile * W, a.mask_i + mrow, a.msf, kBins, u.t0, a.tmax);
        if (BWD && MASK != kStftOut && !de_smem)
            prefetch_rows_l2(threadIdx.x, kTile * W, a.dE + (long long)u.n * n_mels * som, som, n_mels, u.t0, a.tmax);
    };

    int cur_tile = (int)blockIdx.x, cur_ch = 0;              // only these two cross the passes; a unit is decoded where it is used
    if (cur_tile < a.total_tiles) {
        Unit first;
        decode_unit<BWD>(a, cur_tile, 0, first);
        stage(first);
    }
    // the per-CTA tables, under the first unit's copies
    window_fill(S, a.window, threadIdx.x, kTile * W, I16 ? 1.0f / 32768.0f : 1.0f);   // pad column of the scratch <- window table
    if (a.tab_dev) mbar_wait(tab_bar, 0);
    else tables_fill(&sm, tab, threadIdx.x, kTile * W);
    __syncthreads();
#ifdef LMFB_TIMELINE
    unsigned long long gt_loop;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt_loop));
    int n_done = 0;
    long long tl[8];
// after a barrier the clock is only meaningful once something protected by the barrier has been
// touched (BAR.SYNC is deferred-blocking): read a shared word first
#define LMFB_TICK(i) do { volatile float* vs_ = reinterpret_cast<volatile float*>(S); float x_ = vs_[lane]; \
                          asm volatile("" :: "f"(x_) : "memory"); tl[i] = clock64(); } while (0)
#else
#define LMFB_TICK(i) ((void)0)
#endif

    // persistent CTA over (tile, channel) units.  dynamic: the next tile is whichever pending CTA the
    // hardware cancels for us (asked for when a tile starts, read after its pass 1); static:
    // round-robin over the grid.  Either way neighbouring tiles run at the same time, so the sectors
    // they share hit L2.
    while (cur_tile < a.total_tiles) {
        Unit cur;
        decode_unit<BWD>(a, cur_tile, cur_ch, cur);
        if (a.dynamic && (BWD || cur_ch == 0) && threadIdx.x == 0) clc_request(clc_resp, clc_bar);
        const int t = cur.t0 + lane;
        const bool inrow = t < a.tmax;
        const bool valid = t < cur.T;
        const long long row_nm = (long long)cur.n * n_mels * som + t;
        const long long moff = (long long)cur.n * a.msn + (long long)cur.ch * kBins * a.msf + t;
        const bool last_ch = BWD || cur.ch + 1 == nch_fwd;
        Unit nxt;
        if (!BWD && a.frame_lens && cur.t0 == 0 && cur.ch == 0 && threadIdx.x == 0) a.frame_lens[cur.n] = cur.T;

        if (!cur.real) {                            // tile lies entirely in the zero padding
            if (inrow) {
                if (!BWD) {
#pragma unroll 4
                    for (int m = w; m < n_mels; m += W) a.out[row_nm + (unsigned)m * som] = 0.0f;
                } else if (MASK != kMaskNone) {
#pragma unroll 4
                    for (int f = w; f < kBins; f += W) {
                        a.gr[moff + (unsigned)f * a.msf] = 0.0f;
                        if (MASK == kMaskReim || MASK == kStftOut) a.gi[moff + (unsigned)f * a.msf] = 0.0f;
                    }
                }
            }
            int next_tile;
            if (a.dynamic) {
                mbar_wait(clc_bar, clc_phase);
                clc_phase ^= 1u;
                int x;
                next_tile = clc_answer(clc_resp, x) ? x : a.total_tiles;
            } else {
                next_tile = cur.tile + (int)gridDim.x;
            }
            if (next_tile < a.total_tiles) {
                decode_unit<BWD>(a, next_tile, 0, nxt);
                stage(nxt);
            }
            cur_tile = next_tile; cur_ch = 0;
            __syncthreads();                        // (the scheduler's answer is free for the next request)
            continue;
        }

#ifndef LMFB_DBG_NOPREFETCH
        prefetch_unit(cur);                         // pull this unit's mask rows (and dE rows) towards L2 while the FFT runs
#endif
        LMFB_TICK(0);
        // loads of out-of-row lanes are redirected to the last column of the row (always readable)
        const long long clamp = inrow ? 0 : (long long)(a.tmax - 1 - t);
        const float* de = a.dE + row_nm + clamp;
        if (BWD) LMFB_OPAQUE(de);

        if (BWD && de_smem) stage_de<W>(w, lane, de, som_bytes, n_mels, de_buf);
        mbar_wait(raw_bar, raw_phase);              // this unit's rows have landed
        raw_phase ^= 1u;
        LMFB_TICK(1);
        LMFB_TICK(2);
        fft_pass1<W, I16>(w, rawl, col, S + kTile);
        // the row pointers of pass 2 / phase 3, pinned in registers (LMFB_OPAQUE) -- but only from here on: formed
        // ahead of pass 1 they are twelve registers carried through the 64-register FFT, which then spills
        const float* mr = a.mask_r + moff + clamp;
        const float* mi = a.mask_i + moff + clamp;
        float* gr = a.gr + moff;
        float* gi = a.gi + moff;
        float* po = a.out + row_nm;
        LMFB_OPAQUE(mr); LMFB_OPAQUE(mi); LMFB_OPAQUE(gr); LMFB_OPAQUE(gi); LMFB_OPAQUE(po);
        MaskSets<AHEAD> ms;                         // issued before the barrier: the latency hides behind it
        preload_masks<W, MASK, BWD, AHEAD, GW>(w, sm, mr, mi, msf_bytes, ms);
        if constexpr (MASK == kStftOut) {           // no masks: the slot carries the output scale
#pragma unroll
            for (int i = 0; i <= AHEAD; ++i) ms.m[i].vr[0] = valid ? 0.5f : 0.0f;
        }
        LMFB_TICK(3);
        if (BWD && de_smem) cp_async_wait_all();
        __syncthreads();                            // the scratch is complete, the raw buffer is free
        LMFB_TICK(4);

        // which unit comes next, and its rows: they fly under pass 2 / phase 3 of this one
        int next_tile = cur_tile, next_ch = cur_ch + 1;
        if (!last_ch) {
            nxt = cur; nxt.ch = next_ch;
            stage(nxt);
        } else {
            if (a.dynamic) {
                mbar_wait(clc_bar, clc_phase);
                clc_phase ^= 1u;
                int x;
                next_tile = clc_answer(clc_resp, x) ? x : a.total_tiles;
            } else {
                next_tile = cur_tile + (int)gridDim.x;
            }
            next_ch = 0;
            if (next_tile < a.total_tiles) {
                decode_unit<BWD>(a, next_tile, 0, nxt);
                stage(nxt);
            }
        }

        fft_pass2<W, MASK, BWD, AHEAD, GW>(w, col, pl, sm, ms, mr, mi, de, som_bytes, msf_bytes, gr, gi, inrow,
                                           (BWD && de_smem) ? de_buf + lane : nullptr);
        if constexpr (BWD && GW) {                  // gradient into the waveform: adjoint pass 1, overlap-add
            static_assert(!(GW && I16), "an int16 wave takes no gradient");
            StageLane sl;
            stage_lane_init(lane, sl);
            const long long woff = (long long)cur.n * a.wave_stride + (long long)cur.ch * a.wave_stride_ch;
            const int n_rows = (cur.T - cur.t0 < kTile ? cur.T - cur.t0 : kTile) + 1;
            __syncthreads();
            fft_pass1_adj<W>(w, col, S + kTile);
            __syncthreads();
            unstage_tile<W>(w, lane, sl, a.gwave + woff, cur.len, cur.t0, n_rows, S, (a.vec_ok & 2) != 0);
        }
        LMFB_TICK(5);
        if constexpr (!BWD) {
            __syncthreads();
            LMFB_TICK(6);
            if (tab.walkable) {
                phase3_own<W>(w, pl, sm, tab, po, som_bytes, inrow, valid, cur.ch == 0, last_ch);
            } else {
                phase3_gather<W>(w, pl, sm, n_mels, a.mel, po, som_bytes, inrow, valid, cur.ch == 0, last_ch);
            }
        }
#ifdef LMFB_TIMELINE
        else tl[6] = tl[5];
#endif
        LMFB_TICK(7);
#ifdef LMFB_TIMELINE
        if (a.timeline && lane == 0 && blockIdx.x < 64 && n_done < 8) {
            long long* dst = a.timeline + ((long long)n_done * 64 + blockIdx.x) * (W * 8) + w * 8;
            for (int i = 0; i < 8; ++i) dst[i] = tl[i];
        }
        ++n_done;
#endif
        cur_tile = next_tile; cur_ch = next_ch;
        __syncthreads();                            // the scratch (and the scheduler's answer) is free for the next unit
    }
#ifdef LMFB_TIMELINE
    // second view: every CTA's life, stamped with the global nanosecond timer
    if (a.timeline && lane == 0 && w == 0 && blockIdx.x < 1024) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        long long* dst = a.timeline + 8 * 64 * 8 * 8 + (long long)blockIdx.x * 4;
        dst[0] = (long long)gt;                     // end of the CTA (ns)
        dst[1] = (long long)gt_loop;                // start of its tile loop, after the prologue (ns)
        dst[2] = (long long)gt_entry;               // kernel entry of this CTA (ns)
        unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        dst[3] = (long long)smid | ((long long)n_done << 32);
    }
#endif
}

// ------------------------------------------------------------------------------------ K2
// (The per-bin CMVN fused into K1 -- every tile counted with an atomic on its utterance, the block that
// counts the last tile normalising the utterance out of L2 -- was built, is correct, and was measured:
// forward 0.287 ms against 0.203 + 0.019 ms on 256 x 10 s and 58 against 31 + 8 us on 30 x 6 s.  One block
// normalising 40 rows is a serial tail that 1,200 rows spread over the whole GPU do not have, and the
// count's fence and round trip sit on every tile.  The normalisation stays a launch of its own.
// Programmatic dependent launch of the second kernel of a call (CMVN behind K1, K1 behind the CMVN
// gradient, its prologue + staging + pass 1 ahead of griddepcontrol.wait) was measured too, inside the
// CUDA graph of the benchmark: 55.6 vs 54.9 us per step on 30 x 6 s, 0.4329 vs 0.4326 ms on 256 x 10 s:
// nothing, as in round 1.
// Block-per-row kernels with 16-byte accesses (head / aligned float4 body / tail, two float4 per thread for
// 1,001 frames) were measured as well: forward 24.9 vs 20.8 us, backward 37.2 vs 33.1 us on 256 x 10 s -- a
// quarter of the load instructions, but also a quarter of the bytes in flight per SM (4 KB per block of 128
// threads against 16 KB for four warps holding a row each).)
constexpr int kRowThreads = 128;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();                       // protect red[] from the previous use
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}

__device__ __forceinline__ void block_sum2(double& a, double& b, double (*red)[2]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();                       // protect red[] from the previous use
    if ((threadIdx.x & 31) == 0) { red[w][0] = a; red[w][1] = b; }
    __syncthreads();
    a = b = 0.0;
    for (int i = 0; i < nw; ++i) { a += red[i][0]; b += red[i][1]; }
}

__device__ __forceinline__ int frames_of(const int32_t* lengths, int n, int tmax, long long wave_len) {
    int len = lengths[n];
    if (wave_len > 0 && (long long)len > wave_len) len = (int)wave_len;
    int T = len >= 1 ? 1 + len / kHop : 0;
    return T < tmax ? T : tmax;
}

// mode: 1 = per mel bin (grid.x = M rows), 2 = global (grid.x = 1, block walks all M rows)
__global__ void __launch_bounds__(kRowThreads)
cmvn_fwd(float* __restrict__ out, float* __restrict__ stats, const int32_t* __restrict__ lengths,
         int n_mels, int tmax, float eps, int mode, long long wave_len,
         const float* __restrict__ l1_target, float* __restrict__ l1_rows) {
    __shared__ double red[kRowThreads / 32];
    const int n = blockIdx.y;
    const int T = frames_of(lengths, n, tmax, wave_len);
    const int m0 = mode == 1 ? blockIdx.x : 0;
    const int m1 = mode == 1 ? m0 + 1 : n_mels;
    float* base = out + (long long)n * n_mels * tmax;
    const float* tgt = l1_target ? l1_target + (long long)n * n_mels * tmax : nullptr;
    const long long cnt = (long long)(m1 - m0) * T;
    if (T == 0) {                                   // empty utterance: rows are already zero
        for (int m = m0 + threadIdx.x; m < m1; m += kRowThreads) {
            stats[((long long)n * n_mels + m) * 2 + 0] = 0.0f;
            stats[((long long)n * n_mels + m) * 2 + 1] = 1.0f;
        }
        if (tgt) {                                  // (block-uniform) |0 - target| over the whole rows
            for (int m = m0; m < m1; ++m) {
                float a = 0.0f;
                for (int t = threadIdx.x; t < tmax; t += kRowThreads) a += fabsf(tgt[(long long)m * tmax + t]);
                const double tot = block_sum((double)a, red);
                if (threadIdx.x == 0) l1_rows[(long long)n * n_mels + m] = (float)tot;
            }
        }
        return;
    }
    float s = 0.0f;
    for (int m = m0; m < m1; ++m)
        for (int t = threadIdx.x; t < T; t += kRowThreads) s += base[(long long)m * tmax + t];
    const double mean_d = block_sum((double)s, red) / (double)cnt;
    const float mean = (float)mean_d;
    float v = 0.0f;
    for (int m = m0; m < m1; ++m)
        for (int t = threadIdx.x; t < T; t += kRowThreads) {
            const float d = base[(long long)m * tmax + t] - mean;
            v = fmaf(d, d, v);
        }
    const double var = block_sum((double)v, red) / (double)(cnt - 1);
    const float rstd = 1.0f / ((float)sqrt(var) + eps);
    for (int m = m0; m < m1; ++m) {
        float a = 0.0f;
        for (int t = threadIdx.x; t < tmax; t += kRowThreads) {
            const long long i = (long long)m * tmax + t;
            float z = 0.0f;
            if (t < T) { z = (base[i] - mean) * rstd; base[i] = z; }
            if (tgt) a += fabsf(z - tgt[i]);
        }
        if (threadIdx.x == 0) {
            stats[((long long)n * n_mels + m) * 2 + 0] = mean;
            stats[((long long)n * n_mels + m) * 2 + 1] = rstd;
        }
        if (tgt) {                                  // block-uniform
            const double tot = block_sum((double)a, red);
            if (threadIdx.x == 0) l1_rows[(long long)n * n_mels + m] = (float)tot;
        }
    }
}

// dE = dY / (1 + E) with dY the CMVN gradient; mode 0 = no CMVN.
__global__ void __launch_bounds__(kRowThreads)
cmvn_bwd(const float* __restrict__ z, const float* __restrict__ stats,
         const float* __restrict__ grad_out, float* __restrict__ dE,
         const int32_t* __restrict__ lengths, int n_mels, int tmax, float eps, int mode, long long wave_len) {
    __shared__ double red[kRowThreads / 32];
    const int n = blockIdx.y;
    const int T = frames_of(lengths, n, tmax, wave_len);
    const int m0 = mode == 2 ? 0 : blockIdx.x;
    const int m1 = mode == 2 ? n_mels : m0 + 1;
    const long long nb = (long long)n * n_mels * tmax;
    const float* zb = z + nb;
    const float* gb = grad_out + nb;
    float* eb = dE + nb;
    float c1 = 0.0f, k = 0.0f, mean = 0.0f, rstd = 1.0f;
    if (mode != 0) {
        const long long cnt = (long long)(m1 - m0) * T;
        float sg = 0.0f, sgz = 0.0f;
        for (int m = m0; m < m1; ++m)
            for (int t = threadIdx.x; t < T; t += kRowThreads) {
                const long long i = (long long)m * tmax + t;
                const float g = gb[i];
                sg += g;
                sgz = fmaf(g, zb[i], sgz);
            }
        const double sg_d = block_sum((double)sg, red);
        const double sgz_d = block_sum((double)sgz, red);
        mean = stats[((long long)n * n_mels + m0) * 2 + 0];
        rstd = stats[((long long)n * n_mels + m0) * 2 + 1];
        const double sigma = 1.0 / (double)rstd - (double)eps;
        c1 = (float)(sg_d / (double)cnt);
        k = (float)(sgz_d / ((double)(cnt - 1) * sigma));
    }
    for (int m = m0; m < m1; ++m) {
        for (int t = threadIdx.x; t < tmax; t += kRowThreads) {
            const long long i = (long long)m * tmax + t;
            float r = 0.0f;
            if (t < T) {
                const float g = gb[i], zz = zb[i];
                if (mode != 0) {
                    const float dy = fmaf(rstd, g - c1, -zz * k);
                    const float y = fmaf(zz, 1.0f / rstd, mean);
                    r = dy * expf(-y);
                } else {
                    r = g * expf(-zz);
                }
            }
            eb[i] = r;
        }
    }
}

// ---- warp-per-row variants (per-bin CMVN / no CMVN): the row is read ONCE into registers ----
// One warp owns one (utterance, mel) row of up to 32*KMAX frames; no block barrier, no re-read.
constexpr int kRowWarps = 4;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int KMAX>
__global__ void __launch_bounds__(32 * kRowWarps)
cmvn_fwd_rows(float* __restrict__ out, float* __restrict__ stats, const int32_t* __restrict__ lengths,
              int n_mels, int rows, int tmax, float eps, long long wave_len,
              const float* __restrict__ l1_target, float* __restrict__ l1_rows) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int n = row / n_mels;
    const int T = frames_of(lengths, n, tmax, wave_len);
    float* base = out + (long long)row * tmax;
    const float* tgt = l1_target ? l1_target + (long long)row * tmax : nullptr;
    if (T == 0) {
        if (lane == 0) { stats[2 * (long long)row] = 0.0f; stats[2 * (long long)row + 1] = 1.0f; }
        if (tgt) {
            float a = 0.0f;
            for (int t = lane; t < tmax; t += 32) a += fabsf(tgt[t]);
            const double tot = warp_sum((double)a);
            if (lane == 0) l1_rows[row] = (float)tot;
        }
        return;
    }
    float v[KMAX];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int t = lane + 32 * k;
        v[k] = t < T ? base[t] : 0.0f;
        s += v[k];
    }
    const float mean = (float)(warp_sum((double)s) / (double)T);
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const float d = (lane + 32 * k) < T ? v[k] - mean : 0.0f;
        q = fmaf(d, d, q);
    }
    const double var = warp_sum((double)q) / (double)(T - 1);
    const float rstd = 1.0f / ((float)sqrt(var) + eps);
    float a = 0.0f;                                 // optional epilogue: sum |Z - target| of the row (L1Loss_mask,
#pragma unroll                                      // model.py:19-31), while Z is in registers: Z is read once
    for (int k = 0; k < KMAX; ++k) {
        const int t = lane + 32 * k;
        const float z = t < T ? (v[k] - mean) * rstd : 0.0f;
        if (t < T) base[t] = z;
        if (tgt && t < tmax) a += fabsf(z - tgt[t]);
    }
    if (lane == 0) { stats[2 * (long long)row] = mean; stats[2 * (long long)row + 1] = rstd; }
    if (tgt) {                                      // warp-uniform
        const double tot = warp_sum((double)a);
        if (lane == 0) l1_rows[row] = (float)tot;
    }
}

template <int KMAX>
__global__ void __launch_bounds__(32 * kRowWarps)
cmvn_bwd_rows(const float* __restrict__ z, const float* __restrict__ stats,
              const float* __restrict__ grad_out, float* __restrict__ dE,
              const int32_t* __restrict__ lengths, int n_mels, int rows, int tmax, float eps, int mode, long long wave_len) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int n = row / n_mels;
    const int T = frames_of(lengths, n, tmax, wave_len);
    const float* zb = z + (long long)row * tmax;
    const float* gb = grad_out + (long long)row * tmax;
    float* eb = dE + (long long)row * tmax;
    float g[KMAX], zz[KMAX];
    float sg = 0.0f, sgz = 0.0f;
    float c1 = 0.0f, kk = 0.0f, mean = 0.0f, rstd = 1.0f;
    if (mode != 0 && T > 0) {                       // fetched with the row, not behind the reductions
        mean = stats[2 * (long long)row];
        rstd = stats[2 * (long long)row + 1];
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int t = lane + 32 * k;
        g[k] = t < T ? gb[t] : 0.0f;
        zz[k] = t < T ? zb[t] : 0.0f;
        sg += g[k];
        sgz = fmaf(g[k], zz[k], sgz);
    }
    if (mode != 0 && T > 0) {
        const double sg_d = warp_sum((double)sg), sgz_d = warp_sum((double)sgz);
        const double sigma = 1.0 / (double)rstd - (double)eps;
        c1 = (float)(sg_d / (double)T);
        kk = (float)(sgz_d / ((double)(T - 1) * sigma));
    }
    const float inv_rstd = 1.0f / rstd;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int t = lane + 32 * k;
        if (t < tmax) {
            float r = 0.0f;
            if (t < T) {
                if (mode != 0) {
                    const float dy = fmaf(rstd, g[k] - c1, -zz[k] * kk);
                    const float y = fmaf(zz[k], inv_rstd, mean);
                    r = dy * expf(-y);
                } else {
                    r = g[k] * expf(-zz[k]);
                }
            }
            eb[t] = r;
        }
    }
}

// ---- block-per-row forward for rows too long for one warp's registers (1,537 .. 3,072 frames, e.g. every
// 30 s utterance): the row is read ONCE, K elements per thread in registers (the three-pass kernel
// above reads it three times).
template <int K>
__global__ void __launch_bounds__(kRowThreads)
cmvn_fwd_block(float* __restrict__ out, float* __restrict__ stats, const int32_t* __restrict__ lengths,
               int n_mels, int tmax, float eps, long long wave_len,
               const float* __restrict__ l1_target, float* __restrict__ l1_rows) {
    __shared__ double red[kRowThreads / 32];
    const int row = blockIdx.x;
    const int n = row / n_mels;
    const int T = frames_of(lengths, n, tmax, wave_len);
    float* base = out + (long long)row * tmax;
    const float* tgt = l1_target ? l1_target + (long long)row * tmax : nullptr;
    if (T == 0) {                                   // block-uniform
        if (threadIdx.x == 0) { stats[2 * (long long)row] = 0.0f; stats[2 * (long long)row + 1] = 1.0f; }
        if (tgt) {
            float a = 0.0f;
            for (int t = threadIdx.x; t < tmax; t += kRowThreads) a += fabsf(tgt[t]);
            const double tot = block_sum((double)a, red);
            if (threadIdx.x == 0) l1_rows[row] = (float)tot;
        }
        return;
    }
    float v[K];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int t = threadIdx.x + kRowThreads * k;
        v[k] = t < T ? base[t] : 0.0f;
        s += v[k];
    }
    const float mean = (float)(block_sum((double)s, red) / (double)T);
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float d = (threadIdx.x + kRowThreads * k) < T ? v[k] - mean : 0.0f;
        q = fmaf(d, d, q);
    }
    const double var = block_sum((double)q, red) / (double)(T - 1);
    const float rstd = 1.0f / ((float)sqrt(var) + eps);
    float a = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int t = threadIdx.x + kRowThreads * k;
        const float z = t < T ? (v[k] - mean) * rstd : 0.0f;
        if (t < T) base[t] = z;
        if (tgt && t < tmax) a += fabsf(z - tgt[t]);
    }
    if (threadIdx.x == 0) { stats[2 * (long long)row] = mean; stats[2 * (long long)row + 1] = rstd; }
    if (tgt) {                                      // block-uniform
        const double tot = block_sum((double)a, red);
        if (threadIdx.x == 0) l1_rows[row] = (float)tot;
    }
}

// ---- block-per-row backward for rows too long for one warp's registers: the row is read ONCE,
// K elements per thread in registers (the three-pass kernel above reads g and z twice).
template <int K>
__global__ void __launch_bounds__(kRowThreads)
cmvn_bwd_block(const float* __restrict__ z, const float* __restrict__ stats,
               const float* __restrict__ grad_out, float* __restrict__ dE,
               const int32_t* __restrict__ lengths, int n_mels, int tmax, float eps, int mode, long long wave_len) {
    __shared__ double red[kRowThreads / 32][2];
    const int row = blockIdx.x;
    const int n = row / n_mels;
    const int T = frames_of(lengths, n, tmax, wave_len);
    const float* zb = z + (long long)row * tmax;
    const float* gb = grad_out + (long long)row * tmax;
    float* eb = dE + (long long)row * tmax;
    float g[K], zz[K];
    float sg = 0.0f, sgz = 0.0f;
    float c1 = 0.0f, kk = 0.0f, mean = 0.0f, rstd = 1.0f;
    if (mode != 0 && T > 0) {                          // fetched with the row, not behind the reductions
        mean = stats[2 * (long long)row];
        rstd = stats[2 * (long long)row + 1];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int t = threadIdx.x + kRowThreads * k;
        g[k] = t < T ? gb[t] : 0.0f;
        zz[k] = t < T ? zb[t] : 0.0f;
        sg += g[k];
        sgz = fmaf(g[k], zz[k], sgz);
    }
    if (mode != 0) {                                   // block-uniform
        double sg_d = (double)sg, sgz_d = (double)sgz;
        block_sum2(sg_d, sgz_d, red);                  // both sums behind one pair of barriers
        if (T > 0) {
            const double sigma = 1.0 / (double)rstd - (double)eps;
            c1 = (float)(sg_d / (double)T);
            kk = (float)(sgz_d / ((double)(T - 1) * sigma));
        }
    }
    const float inv_rstd = 1.0f / rstd;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int t = threadIdx.x + kRowThreads * k;
        if (t < tmax) {
            float r = 0.0f;
            if (t < T) {
                if (mode != 0) {
                    const float dy = fmaf(rstd, g[k] - c1, -zz[k] * kk);
                    const float y = fmaf(zz[k], inv_rstd, mean);
                    r = dy * expf(-y);
                } else {
                    r = g[k] * expf(-zz[k]);
                }
            }
            eb[t] = r;
        }
    }
}

}  // namespace aas_lmfb

// ======================================================================================
// C ABI
// ======================================================================================
using namespace aas_lmfb;

namespace aas_lmfb {
// generic bases, backward: dP = 1/4 B^T dE as a tensor of kDpRows rows (row f = dP[f]; row 161 = 0),
// which K1 then reads through the identity table.  One thread per (f, t); the basis column is a
// broadcast load, the dE rows are coalesced.  A slow path by design (dense bases are not what the
// trainers use); the banded path never comes here.
__global__ void __launch_bounds__(128)
dp_generic(const float* __restrict__ mel, const float* __restrict__ dE, float* __restrict__ dP, int n_mels, int tmax) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    const int f = blockIdx.y, n = blockIdx.z;
    if (t >= tmax) return;
    float acc = 0.0f;
    if (f < kBins) {
        const float* de = dE + (long long)n * n_mels * tmax + t;
        for (int m = 0; m < n_mels; ++m) acc = fmaf(__ldg(mel + m * kBins + f), de[(long long)m * tmax], acc);
    }
    dP[((long long)n * kDpRows + f) * tmax + t] = 0.25f * acc;
}
}  // namespace aas_lmfb

typedef void (*k1_fwd_fn)(const K1Args, const FwdTab);
typedef void (*k1_bwd_fn)(const K1Args, const BwdTab);

struct K1Variant { int warps, ctas; k1_fwd_fn fwd[3]; k1_bwd_fn bwd[4]; k1_bwd_fn bwd_gw[3]; };
// int16 PCM waves (flag AAS_LMFB_WAVE_I16): the default shape only
static const k1_fwd_fn kFwdI16[3] = { lmfb_k1<kMaskNone, false, 5, 3, false, true>, lmfb_k1<kMaskReim, false, 5, 3, false, true>,
                                      lmfb_k1<kMaskPower, false, 5, 3, false, true> };
static const k1_bwd_fn kBwdI16[4] = { nullptr, lmfb_k1<kMaskReim, true, 5, 3, false, true>, lmfb_k1<kMaskPower, true, 5, 3, false, true>,
                                      lmfb_k1<kStftOut, true, 5, 3, false, true> };   // indexed by mask mode (bwd[3]: STFT output; bwd_gw: with the waveform gradient)

#define LMFB_VARIANT(W, C)                                                                 \
    { W, C,                                                                                \
      { lmfb_k1<kMaskNone, false, W, C>, lmfb_k1<kMaskReim, false, W, C>,                  \
        lmfb_k1<kMaskPower, false, W, C> },                                                \
      { nullptr, lmfb_k1<kMaskReim, true, W, C>, lmfb_k1<kMaskPower, true, W, C>,          \
        lmfb_k1<kStftOut, true, W, C> },                                                   \
      { lmfb_k1<kMaskNone, true, W, C, true>, lmfb_k1<kMaskReim, true, W, C, true>,        \
        lmfb_k1<kMaskPower, true, W, C, true> } }

// (warps per tile, resident CTAs per SM the register budget is sized for).  Shared memory (scratch 42 KB
// + raw buffer 21 KB + tables 3 KB) allows three tiles per SM; five warps per tile give every warp one
// of the five sub-transforms of pass 1.  Launches of at most two rounds of 2 x 148 tiles (config #2: 570
// tiles) are latency- not throughput-bound and take EIGHT warps per tile, two tiles per SM: a tile's
// pass 2 is then 3 steps deep instead of 4 and both rounds are full (measured on 30 x 6 s: forward 26.8
// vs 28.9 us, backward 22.9 vs 27.1 us, step 47.5 vs 54.3 us).
static const K1Variant kVariants[] = {
#ifdef LMFB_ONLY_W5
    LMFB_VARIANT(5, 3), LMFB_VARIANT(4, 3), LMFB_VARIANT(5, 3), LMFB_VARIANT(8, 2),
#else
    LMFB_VARIANT(5, 3), LMFB_VARIANT(4, 3), LMFB_VARIANT(6, 3), LMFB_VARIANT(8, 2),      // (6: measured slower, kept as a tuning shape)
#endif
};
constexpr int kFwdVariantBig = 0, kFwdVariantSmall = 3, kBwdVariantBig = 0, kBwdVariantSmall = 3;
constexpr int kBwdVariantGradWave = 1;
constexpr long long kSmallTiles = 148LL * 2 * 2;            // two rounds of the eight-warp shape

static int variant_of_warps(int warps) {
    for (size_t i = 0; i < sizeof(kVariants) / sizeof(kVariants[0]); ++i)
        if (kVariants[i].warps == warps) return (int)i;
    return -1;
}

struct aas_lmfb_plan {
    FwdTab  fwd;
    BwdTab  bwd;            // banded table, or the identity table over dP when !banded
    int     ml[kBins];      // lower filter of every bin (walkable bases)
    int     n_mels;
    int     banded;         // backward fast path
    alignas(16) unsigned char blob[kTabBlobBytes];   // device image of both shared-memory tables (aas_lmfb_plan_upload)
    int     vfwd, vbwd;     // forced kernel variants (-1: choose by problem size at launch)
    int     static_sched;   // deal tiles round-robin instead of cluster launch control
};

extern "C" int aas_lmfb_abi_version(void) { return AAS_LMFB_ABI_VERSION; }

extern "C" const char* aas_lmfb_strerror(int code) {
    switch (code) {
        case AAS_LMFB_OK:      return "ok";
        case AAS_LMFB_E_NULL:  return "aas_lmfb: required pointer is NULL";
        case AAS_LMFB_E_ALIGN: return "aas_lmfb: buffer is not sufficiently aligned";
        case AAS_LMFB_E_SHAPE: return "aas_lmfb: unsupported shape (n_bins must be 161, 2 <= n_mels <= 128, n >= 0, n_ch >= 1, 1 <= tmax and mask row stride <= 2^22)";
        case AAS_LMFB_E_FLAGS: return "aas_lmfb: invalid mask/cmvn flags";
        case AAS_LMFB_E_MEL:   return "aas_lmfb: this mel basis runs on the generic path, which needs the device copy of the matrix (aas_lmfb_io.mel_dev; use the _ex entry points)";
        case AAS_LMFB_E_NOMEM: return "aas_lmfb: host allocation failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "aas_lmfb: unknown error";
}

extern "C" aas_lmfb_plan* aas_lmfb_plan_create(const float* mel, int n_mels, int n_bins, int* status) {
    int st = AAS_LMFB_OK;
    aas_lmfb_plan* p = nullptr;
    do {
        if (!mel) { st = AAS_LMFB_E_NULL; break; }
        if (n_bins != kBins || n_mels < 2 || n_mels > kMaxMels) { st = AAS_LMFB_E_SHAPE; break; }
        p = new (std::nothrow) aas_lmfb_plan;
        if (!p) { st = AAS_LMFB_E_NOMEM; break; }
        memset(p, 0, sizeof(*p));
        p->n_mels = n_mels;
        p->vfwd = p->vbwd = -1;
        build_fwd_tab(mel, n_mels, &p->fwd, p->ml);
        p->banded = build_bwd_tab(mel, n_mels, &p->bwd) == 0 ? 1 : 0;
        if (!p->banded) build_bwd_tab_identity(&p->bwd);
        tables_image(reinterpret_cast<FwdSmem*>(p->blob), p->fwd);
        tables_image(reinterpret_cast<BwdSmem*>(p->blob + kTabBytesFwd), p->bwd);
    } while (0);
    if (status) *status = st;
    return p;
}

extern "C" void aas_lmfb_plan_destroy(aas_lmfb_plan* plan) { delete plan; }

extern "C" int aas_lmfb_plan_info(const aas_lmfb_plan* plan, int* n_mels, int* fwd_compact, int* bwd_banded) {
    if (!plan) return AAS_LMFB_E_NULL;
    if (n_mels) *n_mels = plan->n_mels;
    if (fwd_compact) *fwd_compact = plan->fwd.walkable;
    if (bwd_banded) *bwd_banded = plan->banded;
    return AAS_LMFB_OK;
}

extern "C" int aas_lmfb_plan_set_tuning(aas_lmfb_plan* plan, int warps_fwd, int warps_bwd, int static_schedule) {
    if (!plan) return AAS_LMFB_E_NULL;
    const int vf = warps_fwd ? variant_of_warps(warps_fwd) : -1, vb = warps_bwd ? variant_of_warps(warps_bwd) : -1;
    if ((warps_fwd && vf < 0) || (warps_bwd && vb < 0)) return AAS_LMFB_E_FLAGS;
    plan->vfwd = vf; plan->vbwd = vb; plan->static_sched = static_schedule ? 1 : 0;
    return AAS_LMFB_OK;
}

extern "C" size_t aas_lmfb_plan_tables_bytes(const aas_lmfb_plan* plan) { return plan ? (size_t)kTabBlobBytes : 0; }

extern "C" int aas_lmfb_plan_upload(const aas_lmfb_plan* plan, void* tables_dev, void* cuda_stream) {
    if (!plan || !tables_dev) return AAS_LMFB_E_NULL;
    if ((uintptr_t)tables_dev & 15u) return AAS_LMFB_E_ALIGN;
    return (int)cudaMemcpyAsync(tables_dev, plan->blob, kTabBlobBytes, cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream);
}

static size_t round16(size_t b) { return (b + 15) & ~(size_t)15; }

extern "C" size_t aas_lmfb_workspace_bytes(const aas_lmfb_plan* plan, int n, int tmax, uint32_t /*flags*/) {
    if (!plan || n <= 0 || tmax <= 0) return 0;
    size_t b = round16((size_t)n * (size_t)plan->n_mels * (size_t)tmax * sizeof(float));       // dE
    if (!plan->banded) b += round16((size_t)n * (size_t)kDpRows * (size_t)tmax * sizeof(float));   // dP
    return b;
}

namespace {

// cudaFuncSetAttribute is per (function, device); do it once each so that launches inside
// a CUDA-graph capture are pure stream work.
int ensure_attrs(const void* fn, int smem) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int> > done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<const void*, int> key(fn, dev);
    if (done.count(key)) return 0;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return (int)e;
    done.insert(key);
    return 0;
}

// the calling thread's device for the duration of a call (autograd runs backward on another host
// thread than forward: nothing here may depend on thread-local state set elsewhere)
struct DeviceScope {
    int prev = -1, rc = 0;
    explicit DeviceScope(int want) {
        if (want < 0) return;
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) { rc = (int)e; prev = -1; return; }
        if (prev == want) { prev = -1; return; }
        e = cudaSetDevice(want);
        if (e != cudaSuccess) { rc = (int)e; prev = -1; }
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

template <class Fn, class Tab>
int launch_k1(const aas_lmfb_plan* plan, const K1Variant& v, Fn fn, K1Args& a, const Tab& tab, bool bwd,
              long long units, cudaStream_t stream) {
    if (!fn) return AAS_LMFB_E_FLAGS;
    const int smem = smem_bytes(bwd);
    const int rc = ensure_attrs((const void*)fn, smem);
    if (rc) return rc;
    const long long total = units * a.tiles_per_utt;               // units: utterances (forward) or utterance-channels
    if (total <= 0) return AAS_LMFB_OK;
    if (total > 0x7fffffffLL) return AAS_LMFB_E_SHAPE;
    a.total_tiles = (int)total;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    int per_sm = 233472 / (smem + 1024);                           // shared memory per SM / per-CTA footprint
    if (per_sm > v.ctas) per_sm = v.ctas;
    if (per_sm < 1) per_sm = 1;
    const long long resident = (long long)sms * per_sm;           // one persistent CTA per scratch slot
    // more tiles than resident CTAs: one CTA per tile, the resident ones take over the pending ones
    a.dynamic = (!plan->static_sched && total > resident) ? 1 : 0;
    const unsigned blocks = (unsigned)(a.dynamic || total < resident ? total : resident);
    fn<<<blocks, kTile * v.warps, smem, stream>>>(a, tab);
    return (int)cudaPeekAtLastError();
}

int check_common(const aas_lmfb_plan* plan, const aas_lmfb_io* io) {
    if (!plan || !io) return AAS_LMFB_E_NULL;
    if (io->struct_size < sizeof(aas_lmfb_io)) return AAS_LMFB_E_SHAPE;
    if (!io->window || (io->n > 0 && (!io->wave || !io->lengths))) return AAS_LMFB_E_NULL;
    if (io->n < 0 || io->n_ch < 1 || io->tmax < 1 || io->tmax > (1 << 22)) return AAS_LMFB_E_SHAPE;
    const unsigned mask = io->flags & 3u, cm = (io->flags >> 2) & 3u;
    if (mask > 2u || cm > 2u || (io->flags >> 5)) return AAS_LMFB_E_FLAGS;
    if (mask != AAS_LMFB_MASK_NONE && !io->mask_r) return AAS_LMFB_E_NULL;
    if (mask == AAS_LMFB_MASK_REIM && !io->mask_i) return AAS_LMFB_E_NULL;
    if (mask != AAS_LMFB_MASK_NONE && (io->mask_stride_f < io->tmax || io->mask_stride_f > (1 << 22))) return AAS_LMFB_E_SHAPE;
    const bool i16 = (io->flags & AAS_LMFB_WAVE_I16) != 0;
    const uintptr_t al = (uintptr_t)io->mask_r | (uintptr_t)io->mask_i | (uintptr_t)io->window | (uintptr_t)io->mel_dev;
    if ((al & 3u) || ((uintptr_t)io->wave & (i16 ? 1u : 3u)) || ((uintptr_t)io->tables & 15u)) return AAS_LMFB_E_ALIGN;
    return AAS_LMFB_OK;
}

void fill_args(K1Args& a, const aas_lmfb_io* io) {
    memset(&a, 0, sizeof(a));
    a.wave = io->wave; a.lengths = io->lengths; a.wave_stride = io->wave_stride;
    a.wave_stride_ch = io->wave_stride_ch; a.wave_len = io->wave_len; a.n_ch = io->n_ch;
    a.mask_r = io->mask_r; a.mask_i = io->mask_i; a.msn = io->mask_stride_n; a.msf = (unsigned)io->mask_stride_f;
    a.window = io->window; a.tmax = io->tmax; a.mel = io->mel_dev;
    a.tab_dev = nullptr;
    a.tiles_per_utt = (io->tmax + kTile - 1) / kTile;
    a.div_tpu = make_fastdiv(a.tiles_per_utt);
    a.div_nch = make_fastdiv(io->n_ch);
    // bit 0: every row start is 16-byte aligned (bulk copies); bit 1: 8-byte aligned (vector atomics of the adjoint staging)
    const int per16 = (io->flags & AAS_LMFB_WAVE_I16) ? 8 : 4;          // samples per 16 bytes
    a.vec_ok = ((((uintptr_t)io->wave & 15u) == 0 && io->wave_stride % per16 == 0 && io->wave_stride_ch % per16 == 0) ? 1 : 0) |
               ((((uintptr_t)io->wave & 7u) == 0 && (io->wave_stride & 1) == 0 && (io->wave_stride_ch & 1) == 0) ? 2 : 0);
}

void rec(void* const* prof, int i, cudaStream_t s) {
    if (prof && prof[i]) cudaEventRecord((cudaEvent_t)prof[i], s);
}

}  // namespace

extern "C" int aas_lmfb_forward_ex(const aas_lmfb_plan* plan, const aas_lmfb_io* io) {
    int rc = check_common(plan, io);
    if (rc) return rc;
    const int n = io->n, tmax = io->tmax;
    if (n == 0) return AAS_LMFB_OK;
    const unsigned mask = io->flags & 3u, cm = (io->flags >> 2) & 3u;
    float* out = io->out; float* stats = io->stats;
    if (!out || (cm != 0 && !stats)) return AAS_LMFB_E_NULL;
    if (((uintptr_t)out | (uintptr_t)stats) & 3u) return AAS_LMFB_E_ALIGN;
    if (!plan->fwd.walkable && !io->mel_dev) return AAS_LMFB_E_MEL;
    if ((io->l1_target != nullptr) != (io->l1_rows != nullptr)) return AAS_LMFB_E_NULL;
    if (io->l1_target && cm == 0) return AAS_LMFB_E_FLAGS;          // the epilogue lives in the CMVN kernel
    if (((uintptr_t)io->l1_target | (uintptr_t)io->l1_rows) & 3u) return AAS_LMFB_E_ALIGN;
    DeviceScope scope(io->device);
    if (scope.rc) return scope.rc;
    cudaStream_t stream = (cudaStream_t)io->cuda_stream;
    void* const* prof = io->prof;

    K1Args a;
    fill_args(a, io);
    a.out = out;
    a.tab_dev = io->tables;
    a.frame_lens = io->frame_lens;
#ifdef LMFB_TIMELINE
    { const char* e = getenv("AAS_LMFB_TIMELINE_FWD"); a.timeline = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
#endif
    const bool small = (long long)n * a.tiles_per_utt <= kSmallTiles;
    const bool i16 = (io->flags & AAS_LMFB_WAVE_I16) != 0;
    const K1Variant& v = kVariants[i16 ? 0 : (plan->vfwd >= 0 ? plan->vfwd : (small ? kFwdVariantSmall : kFwdVariantBig))];
    FwdTab band = plan->fwd;
    set_warp_ranges(&band, plan->ml, v.warps);
    rec(prof, 0, stream);
    rc = (io->flags & AAS_LMFB_WAVE_I16) ? launch_k1(plan, v, kFwdI16[mask], a, band, false, n, stream)
                                         : launch_k1(plan, v, v.fwd[mask], a, band, false, n, stream);
    rec(prof, 1, stream);
    if (rc) return rc;
    rec(prof, 2, stream);
    if (cm != 0) {
        const int rows = n * plan->n_mels;
        const float eps = io->eps;
        const int32_t* lengths = io->lengths;
        const unsigned blocks = (unsigned)((rows + kRowWarps - 1) / kRowWarps);
        if (cm == 1 && tmax <= 32 * 8) {
            cmvn_fwd_rows<8><<<blocks, 32 * kRowWarps, 0, stream>>>(out, stats, lengths, plan->n_mels, rows, tmax, eps, io->wave_len, io->l1_target, io->l1_rows);
        } else if (cm == 1 && tmax <= 32 * 24) {
            cmvn_fwd_rows<24><<<blocks, 32 * kRowWarps, 0, stream>>>(out, stats, lengths, plan->n_mels, rows, tmax, eps, io->wave_len, io->l1_target, io->l1_rows);
        } else if (cm == 1 && tmax <= 32 * 32) {
            cmvn_fwd_rows<32><<<blocks, 32 * kRowWarps, 0, stream>>>(out, stats, lengths, plan->n_mels, rows, tmax, eps, io->wave_len, io->l1_target, io->l1_rows);
        } else if (cm == 1 && tmax <= 32 * 48) {
            cmvn_fwd_rows<48><<<blocks, 32 * kRowWarps, 0, stream>>>(out, stats, lengths, plan->n_mels, rows, tmax, eps, io->wave_len, io->l1_target, io->l1_rows);
        } else if (cm == 1 && tmax <= kRowThreads * 24) {
            cmvn_fwd_block<24><<<(unsigned)rows, kRowThreads, 0, stream>>>(out, stats, lengths, plan->n_mels, tmax, eps, io->wave_len, io->l1_target, io->l1_rows);
        } else {
            dim3 grid(cm == 1 ? plan->n_mels : 1, n);
            cmvn_fwd<<<grid, kRowThreads, 0, stream>>>(out, stats, lengths, plan->n_mels, tmax, eps, (int)cm, io->wave_len, io->l1_target, io->l1_rows);
        }
        rc = (int)cudaPeekAtLastError();
    }
    rec(prof, 3, stream);
    return rc;
}

extern "C" int aas_lmfb_backward_ex(const aas_lmfb_plan* plan, const aas_lmfb_io* io) {
    int rc = check_common(plan, io);
    if (rc) return rc;
    const int n = io->n, tmax = io->tmax;
    if (n == 0) return AAS_LMFB_OK;
    const unsigned mask = io->flags & 3u, cm = (io->flags >> 2) & 3u;
    float* grad_wave = io->grad_wave;
    if (mask == AAS_LMFB_MASK_NONE && !grad_wave) return AAS_LMFB_E_FLAGS;        // nothing to differentiate into
    const bool i16 = (io->flags & AAS_LMFB_WAVE_I16) != 0;
    if (i16 && grad_wave) return AAS_LMFB_E_FLAGS;                                // integer samples take no gradient
    if (!io->out || !io->grad_out || !io->workspace || (mask != AAS_LMFB_MASK_NONE && !io->grad_mask_r) || (cm != 0 && !io->stats)) return AAS_LMFB_E_NULL;
    if ((uintptr_t)grad_wave & 7u) return AAS_LMFB_E_ALIGN;
    if (mask == AAS_LMFB_MASK_REIM && !io->grad_mask_i) return AAS_LMFB_E_NULL;
    if (((uintptr_t)io->out | (uintptr_t)io->grad_out | (uintptr_t)io->workspace | (uintptr_t)io->grad_mask_r |
         (uintptr_t)io->grad_mask_i | (uintptr_t)io->stats) & 3u) return AAS_LMFB_E_ALIGN;
    if (!plan->banded && !io->mel_dev) return AAS_LMFB_E_MEL;
    // gradient rows of different utterances / channels must not overlap (they are written independently)
    if (mask != AAS_LMFB_MASK_NONE && n > 1 && io->mask_stride_n < (int64_t)io->n_ch * kBins * io->mask_stride_f) return AAS_LMFB_E_SHAPE;
    DeviceScope scope(io->device);
    if (scope.rc) return scope.rc;
    cudaStream_t stream = (cudaStream_t)io->cuda_stream;
    void* const* prof = io->prof;
    float* dE = (float*)io->workspace;
    const float eps = io->eps;
    const int32_t* lengths = io->lengths;
    const float* out = io->out; const float* stats = io->stats; const float* grad_out = io->grad_out;

    rec(prof, 2, stream);
    {
        const int rows = n * plan->n_mels;
        const unsigned blocks = (unsigned)((rows + kRowWarps - 1) / kRowWarps);
        if (cm != 2 && tmax <= 32 * 8) {
            cmvn_bwd_rows<8><<<blocks, 32 * kRowWarps, 0, stream>>>(out, stats, grad_out, dE, lengths, plan->n_mels, rows, tmax, eps, (int)cm, io->wave_len);
        } else if (cm != 2 && tmax <= 32 * 24) {
            cmvn_bwd_rows<24><<<blocks, 32 * kRowWarps, 0, stream>>>(out, stats, grad_out, dE, lengths, plan->n_mels, rows, tmax, eps, (int)cm, io->wave_len);
        } else if (cm != 2 && tmax <= kRowThreads * 8) {
            cmvn_bwd_block<8><<<(unsigned)rows, kRowThreads, 0, stream>>>(out, stats, grad_out, dE, lengths, plan->n_mels, tmax, eps, (int)cm, io->wave_len);
        } else if (cm != 2 && tmax <= kRowThreads * 24) {
            cmvn_bwd_block<24><<<(unsigned)rows, kRowThreads, 0, stream>>>(out, stats, grad_out, dE, lengths, plan->n_mels, tmax, eps, (int)cm, io->wave_len);
        } else {
            dim3 grid(cm == 2 ? 1 : plan->n_mels, n);
            cmvn_bwd<<<grid, kRowThreads, 0, stream>>>(out, stats, grad_out, dE, lengths, plan->n_mels, tmax, eps, (int)cm, io->wave_len);
        }
        rc = (int)cudaPeekAtLastError();
        if (rc == 0 && !plan->banded) {                          // generic basis: dP = 1/4 B^T dE, read through the identity table
            float* dP = (float*)((char*)io->workspace + round16((size_t)n * plan->n_mels * tmax * sizeof(float)));
            dim3 grid((unsigned)((tmax + 127) / 128), kDpRows, n);
            dp_generic<<<grid, 128, 0, stream>>>(io->mel_dev, dE, dP, plan->n_mels, tmax);
            rc = (int)cudaPeekAtLastError();
            dE = dP;
        }
    }
    rec(prof, 3, stream);
    if (rc) return rc;

    K1Args a;
    fill_args(a, io);
    a.dE = dE; a.gr = io->grad_mask_r; a.gi = io->grad_mask_i;
    a.tab_dev = io->tables ? (const char*)io->tables + kTabBytesFwd : nullptr;
#ifdef LMFB_TIMELINE
    { const char* e = getenv("AAS_LMFB_TIMELINE_BWD"); a.timeline = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
#endif
    const long long units = (long long)n * io->n_ch;
    const bool small = units * a.tiles_per_utt <= kSmallTiles;
    // (with the waveform gradient: four warps per tile -- 168 registers, no spills; measured 0.65 vs 0.73 ms on 256 x 10 s)
    const K1Variant& v = kVariants[i16 ? 0 : (plan->vbwd >= 0 ? plan->vbwd : grad_wave ? kBwdVariantGradWave
                                                                 : (small ? kBwdVariantSmall : kBwdVariantBig))];
    if (mask == AAS_LMFB_MASK_NONE) { a.mask_r = a.mask_i = (const float*)io->wave; a.gr = a.gi = nullptr; a.msf = (unsigned)tmax; a.msn = 0; }   // never dereferenced
    a.gwave = grad_wave;
    if (grad_wave) {                                             // the kernel ADDS (overlapping frames, reflect padding)
        const int64_t row = io->n_ch > 1 ? io->wave_stride_ch : io->wave_stride;       // rows are (n, ch) pairs
        int64_t w64 = row < (int64_t)tmax * kHop ? row : (int64_t)tmax * kHop;
        if (io->wave_len > 0 && io->wave_len < w64) w64 = io->wave_len;            // never past the end of a row
        const size_t width = (size_t)w64;
        cudaError_t e = cudaSuccess;
        if (io->n_ch == 1 || io->wave_stride == io->wave_stride_ch * io->n_ch) {
            e = cudaMemset2DAsync(grad_wave, (size_t)row * sizeof(float), 0, width * sizeof(float), (size_t)units, stream);
        } else {
            for (int i = 0; i < n && e == cudaSuccess; ++i)
                e = cudaMemset2DAsync(grad_wave + (size_t)i * io->wave_stride, (size_t)row * sizeof(float), 0,
                                      width * sizeof(float), (size_t)io->n_ch, stream);
        }
        if (e != cudaSuccess) return (int)e;
    }
    rec(prof, 0, stream);
    rc = grad_wave ? launch_k1(plan, v, v.bwd_gw[mask], a, plan->bwd, true, units, stream)
         : i16     ? launch_k1(plan, v, kBwdI16[mask], a, plan->bwd, true, units, stream)
                   : launch_k1(plan, v, v.bwd[mask], a, plan->bwd, true, units, stream);
    rec(prof, 1, stream);
    return rc;
}

// ---- classic entry points: one channel, the calling thread's device -------------------------------
static void classic_io(aas_lmfb_io& io, const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                       const float* mask_r, const float* mask_i, int64_t msn, int64_t msf, const float* window,
                       int tmax, uint32_t flags, float eps, void* stream, void* const* prof) {
    memset(&io, 0, sizeof(io));
    io.struct_size = (uint32_t)sizeof(io); io.flags = flags; io.device = -1; io.n = n; io.n_ch = 1; io.tmax = tmax;
    io.eps = eps; io.wave = wave; io.wave_stride = wave_stride; io.wave_stride_ch = 0; io.wave_len = 0;
    io.lengths = lengths; io.mask_r = mask_r; io.mask_i = mask_i; io.mask_stride_n = msn; io.mask_stride_f = msf;
    io.window = window; io.cuda_stream = stream; io.prof = prof;
}

extern "C" int aas_lmfb_forward(const aas_lmfb_plan* plan,
                                const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                                const float* mask_r, const float* mask_i,
                                int64_t mask_stride_n, int64_t mask_stride_f,
                                const float* window,
                                float* out, float* stats, int tmax,
                                uint32_t flags, float eps, void* cuda_stream, void* const* prof) {
    aas_lmfb_io io;
    classic_io(io, wave, lengths, n, wave_stride, mask_r, mask_i, mask_stride_n, mask_stride_f, window, tmax, flags, eps, cuda_stream, prof);
    io.out = out; io.stats = stats;
    return aas_lmfb_forward_ex(plan, &io);
}

extern "C" int aas_lmfb_backward(const aas_lmfb_plan* plan,
                                 const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                                 const float* mask_r, const float* mask_i,
                                 int64_t mask_stride_n, int64_t mask_stride_f,
                                 const float* window,
                                 const float* out, const float* stats, const float* grad_out,
                                 float* grad_mask_r, float* grad_mask_i,
                                 void* workspace, int tmax,
                                 uint32_t flags, float eps, void* cuda_stream, void* const* prof) {
    aas_lmfb_io io;
    classic_io(io, wave, lengths, n, wave_stride, mask_r, mask_i, mask_stride_n, mask_stride_f, window, tmax, flags, eps, cuda_stream, prof);
    io.out = const_cast<float*>(out); io.stats = const_cast<float*>(stats); io.grad_out = grad_out;
    io.grad_mask_r = grad_mask_r; io.grad_mask_i = grad_mask_i; io.workspace = workspace;
    return aas_lmfb_backward_ex(plan, &io);
}

extern "C" int aas_lmfb_backward_wave(const aas_lmfb_plan* plan,
                                      const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                                      const float* mask_r, const float* mask_i,
                                      int64_t mask_stride_n, int64_t mask_stride_f,
                                      const float* window,
                                      const float* out, const float* stats, const float* grad_out,
                                      float* grad_mask_r, float* grad_mask_i, float* grad_wave,
                                      void* workspace, int tmax,
                                      uint32_t flags, float eps, void* cuda_stream) {
    if (!grad_wave) return AAS_LMFB_E_NULL;
    aas_lmfb_io io;
    classic_io(io, wave, lengths, n, wave_stride, mask_r, mask_i, mask_stride_n, mask_stride_f, window, tmax, flags, eps, cuda_stream, nullptr);
    io.out = const_cast<float*>(out); io.stats = const_cast<float*>(stats); io.grad_out = grad_out;
    io.grad_mask_r = grad_mask_r; io.grad_mask_i = grad_mask_i; io.grad_wave = grad_wave; io.workspace = workspace;
    return aas_lmfb_backward_ex(plan, &io);
}

// STFT as an output: what BRNNmultiCH.forward takes as its input, (N, 2*F, T) with the real rows first
// and the imaginary rows second (model.py:170, :186-188).  Same staging + FFT as the other kernels;
// the spectrum rows are stored straight from pass 2.
extern "C" int aas_lmfb_stft(const aas_lmfb_plan* plan,
                             const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                             const float* window, float* out, int64_t out_stride_n, int tmax,
                             void* cuda_stream) {
    aas_lmfb_io io;
    classic_io(io, wave, lengths, n, wave_stride, nullptr, nullptr, 0, 0, window, tmax, AAS_LMFB_MASK_NONE, 0.0f, cuda_stream, nullptr);
    int rc = check_common(plan, &io);
    if (rc) return rc;
    if (n == 0) return AAS_LMFB_OK;
    if (!out) return AAS_LMFB_E_NULL;
    if ((uintptr_t)out & 3u) return AAS_LMFB_E_ALIGN;
    if (out_stride_n < 2LL * kBins * tmax) return AAS_LMFB_E_SHAPE;
    K1Args a;
    fill_args(a, &io);
    a.mask_r = a.mask_i = out;                                    // never read (clamped pointer arithmetic only)
    a.msn = out_stride_n; a.msf = (unsigned)tmax;
    a.dE = out; a.gr = out; a.gi = out + (long long)kBins * tmax;
    const bool small = (long long)n * a.tiles_per_utt <= kSmallTiles;
    const K1Variant& v = kVariants[plan->vbwd >= 0 ? plan->vbwd : (small ? kBwdVariantSmall : kBwdVariantBig)];
    return launch_k1(plan, v, v.bwd[3], a, plan->bwd, true, n, (cudaStream_t)cuda_stream);
}

// ======================================================================================
// L1Loss_mask (Speech_enhancement_by_AAS/model.py:19-31): the loss every trainer applies to the
// LMFB features right after this front-end (trainer_AAS.py:146-161, :176-181; trainer_DCE.py).
//   loss = sum |input - target| / nElement,   nElement = numel(mask) - sum(mask)  (unmasked FRAMES)
// The reference's masked_fill is not in-place, i.e. a no-op: padded frames DO contribute.  That
// behaviour is reproduced when `mask` is NULL; passing the byte mask zeroes the padded frames
// (the evident intent), as an opt-in.
// ======================================================================================
namespace aas_lmfb {

constexpr int kL1Threads = 256;
constexpr int kL1MaxBlocks = 1184;           // 148 SMs x 8

__global__ void __launch_bounds__(kL1Threads)
l1_abs_partial(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ mask,
               long long total, int c, int tmax, float* __restrict__ partial) {
    __shared__ double red[kL1Threads / 32];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * kL1Threads + threadIdx.x; i < total;
         i += (long long)gridDim.x * kL1Threads) {
        float d = fabsf(a[i] - b[i]);
        if (mask) {
            const long long n = i / ((long long)c * tmax);
            const int t = (int)(i % tmax);
            if (mask[n * tmax + t]) d = 0.0f;
        }
        s += (double)d;
    }
    const double tot = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = (float)tot;
}

__global__ void __launch_bounds__(kL1Threads)
l1_abs_final(const float* __restrict__ partial, int count, float* __restrict__ out) {
    __shared__ double red[kL1Threads / 32];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += kL1Threads) s += (double)partial[i];   // fixed order: deterministic
    const double tot = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = (float)tot;
}

__global__ void __launch_bounds__(kL1Threads)
l1_abs_grad(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ mask,
            long long total, int c, int tmax, const float* __restrict__ scale,
            float* __restrict__ ga, float* __restrict__ gb) {
    const float sc = scale[0];
    for (long long i = (long long)blockIdx.x * kL1Threads + threadIdx.x; i < total;
         i += (long long)gridDim.x * kL1Threads) {
        const float d = a[i] - b[i];
        float g = d > 0.0f ? sc : (d < 0.0f ? -sc : 0.0f);            // torch: sign(0) = 0
        if (mask) {
            const long long n = i / ((long long)c * tmax);
            const int t = (int)(i % tmax);
            if (mask[n * tmax + t]) g = 0.0f;
        }
        if (ga) ga[i] = g;
        if (gb) gb[i] = -g;
    }
}

}  // namespace aas_lmfb

extern "C" int aas_l1_partial_count(void) { return kL1MaxBlocks; }

extern "C" int aas_l1_abs_sum(const float* a, const float* b, const uint8_t* mask, int n, int c, int tmax,
                              float* partial, float* out, void* cuda_stream) {
    if (!a || !b || !partial || !out) return AAS_LMFB_E_NULL;
    if (n < 0 || c < 1 || tmax < 1) return AAS_LMFB_E_SHAPE;
    const long long total = (long long)n * c * tmax;
    long long blocks = (total + kL1Threads - 1) / kL1Threads;
    if (blocks > kL1MaxBlocks) blocks = kL1MaxBlocks;
    if (blocks < 1) blocks = 1;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    l1_abs_partial<<<(unsigned)blocks, kL1Threads, 0, stream>>>(a, b, mask, total, c, tmax, partial);
    l1_abs_final<<<1, kL1Threads, 0, stream>>>(partial, (int)blocks, out);
    return (int)cudaPeekAtLastError();
}

// sum of the per-row partial sums the forward's L1 epilogue wrote (aas_lmfb_io.l1_rows), in a fixed order
extern "C" int aas_l1_rows_sum(const float* rows, int count, float* out, void* cuda_stream) {
    if (!rows || !out) return AAS_LMFB_E_NULL;
    if (count < 0) return AAS_LMFB_E_SHAPE;
    l1_abs_final<<<1, kL1Threads, 0, (cudaStream_t)cuda_stream>>>(rows, count, out);
    return (int)cudaPeekAtLastError();
}

extern "C" int aas_l1_abs_grad(const float* a, const float* b, const uint8_t* mask, int n, int c, int tmax,
                               const float* scale, float* grad_a, float* grad_b, void* cuda_stream) {
    if (!a || !b || !scale || (!grad_a && !grad_b)) return AAS_LMFB_E_NULL;
    if (n < 0 || c < 1 || tmax < 1) return AAS_LMFB_E_SHAPE;
    const long long total = (long long)n * c * tmax;
    if (total == 0) return AAS_LMFB_OK;
    long long blocks = (total + kL1Threads - 1) / kL1Threads;
    if (blocks > kL1MaxBlocks) blocks = kL1MaxBlocks;
    l1_abs_grad<<<(unsigned)blocks, kL1Threads, 0, (cudaStream_t)cuda_stream>>>(a, b, mask, total, c, tmax, scale,
                                                                               grad_a, grad_b);
    return (int)cudaPeekAtLastError();
}

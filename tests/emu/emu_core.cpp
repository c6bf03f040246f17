// CPU emulation of the per-thread device code in aas_enhancement_b200/csrc/lmfb_core.cuh.
// TEST INFRASTRUCTURE ONLY: lets the index algebra of the lane=frame FFT be checked on a
// machine without a GPU.  A CTA of W warps is emulated phase by phase (a block barrier becomes
// "finish the phase for every warp and lane before starting the next one").
#include <vector>
#include <cstring>
#include "lmfb_core.cuh"
#include "mel_band.hpp"

using namespace aas_lmfb;

namespace {

template <int W, int MASK, bool BWD>
void emu_tile(const float* wave_row, int len, int t0, const float* window, bool vec_ok,
              const typename TabOf<BWD>::Param& tab, const float* mr, const float* mi, unsigned msf,
              const float* dE, float* out, unsigned som, float* gr, float* gi, int tmax, int T,
              std::vector<float2>& S, float* gwave_row = nullptr) {
    typename TabOf<BWD>::Smem sm;
    tables_fill(&sm, tab, 0, 1);
    window_fill(S.data(), window, 0, 1);
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane) {
            StageLane sl;
            stage_lane_init(lane, S.data(), sl);
            const int n_rows = (T - t0 < kTile ? T - t0 : kTile) + 1;
            stage_tile<W>(w, lane, sl, wave_row, len, t0, n_rows, S.data(), vec_ok);
        }
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane) fft_pass1<W>(w, S.data() + lane, S.data() + kTile);
    // pass 2: a real warp runs its lanes in lockstep (all loads of a step before its stores); the
    // emulation runs lane after lane, which is equivalent only if no lane's store can hit a word
    // another lane of the same step still has to read.  The compact P rows overlap the float2
    // words of OTHER lanes (same slot), so emulate the lockstep by double-buffering the scratch.
    std::vector<float2> Sin(S);
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane) {
            const int t = t0 + lane;
            const bool inrow = t < tmax;
            const long long clamp = inrow ? 0 : (long long)(tmax - 1 - t);
            const float* mrp = mr ? mr + t + clamp : nullptr;
            const float* mip = mi ? mi + t + clamp : nullptr;
            const float* dep = dE ? dE + t + clamp : nullptr;
            constexpr int AHEAD = BWD ? kAheadBwd : kAheadFwd + 1;   // (the forward emulation keeps exercising the three-set rotation)
            MaskSets<AHEAD> ms;
            if (BWD && gwave_row) preload_masks<W, MASK, BWD, AHEAD, true>(w, sm, mrp, mip, msf * 4u, ms);
            else                  preload_masks<W, MASK, BWD, AHEAD, false>(w, sm, mrp, mip, msf * 4u, ms);
            if (BWD && gwave_row)
                fft_pass2<W, MASK, BWD, AHEAD, true>(w, Sin.data() + lane, reinterpret_cast<float*>(S.data()) + lane, sm, ms,
                                        mrp, mip, dep, som * 4u, msf * 4u,
                                        gr ? gr + t : nullptr, gi ? gi + t : nullptr, inrow);
            else
                fft_pass2<W, MASK, BWD, AHEAD, false>(w, Sin.data() + lane, reinterpret_cast<float*>(S.data()) + lane, sm, ms,
                                        mrp, mip, dep, som * 4u, msf * 4u,
                                        gr ? gr + t : nullptr, gi ? gi + t : nullptr, inrow);
        }
    if (BWD && gwave_row) {                                // gradient into the waveform (adjoint pass 1, overlap-add)
        for (int w = 0; w < W; ++w)
            for (int lane = 0; lane < 32; ++lane) fft_pass1_adj<W>(w, Sin.data() + lane, Sin.data() + kTile);
        for (int w = 0; w < W; ++w)
            for (int lane = 0; lane < 32; ++lane) {
                StageLane sl;
                stage_lane_init(lane, Sin.data(), sl);
                const int n_rows = (T - t0 < kTile ? T - t0 : kTile) + 1;
                unstage_tile<W>(w, lane, sl, gwave_row, len, t0, n_rows, Sin.data(), vec_ok);
            }
    }
    if constexpr (!BWD) {
        for (int w = 0; w < W; ++w)
            for (int lane = 0; lane < 32; ++lane)
                phase3_walk<W>(w, reinterpret_cast<float*>(S.data()) + lane, sm, tab);
        for (int w = 0; w < W; ++w)
            for (int lane = 0; lane < 32; ++lane) {
                const int t = t0 + lane;
                phase3_finish<W>(w, reinterpret_cast<float*>(S.data()) + lane, tab, out + t, som * 4u, t < tmax, t < T);
            }
    }
}

template <int W, int MASK>
void emu_k1_impl(int bwd, const float* wave, const int* lengths, int n_utt, long long wave_stride,
                 const float* mask_r, const float* mask_i, long long msn, long long msf,
                 const float* window, const FwdTab& ft, const BwdTab& bt, float* out, const float* dE,
                 float* gr, float* gi, int tmax, int vec_ok, float* gwave = nullptr) {
    const int tiles = (tmax + kTile - 1) / kTile;
    const int n_mels = ft.n_mels;
    std::vector<float2> S(kSlots * kPitch);
    for (int n = 0; n < n_utt; ++n)
        for (int tile = 0; tile < tiles; ++tile) {
            const int t0 = tile * kTile;
            const int len = lengths[n];
            int T = len >= 1 ? 1 + len / kHop : 0;
            T = T < tmax ? T : tmax;
            const unsigned som = (unsigned)tmax;
            const long long nb = (long long)n * n_mels * som;
            if (t0 >= T) {
                for (int lane = 0; lane < 32; ++lane) {
                    const int t = t0 + lane;
                    if (t >= tmax) continue;
                    if (!bwd) for (int m = 0; m < n_mels; ++m) out[nb + t + (long long)m * som] = 0.0f;
                    else if (MASK != kMaskNone) for (int f = 0; f < kBins; ++f) {
                        gr[(long long)n * msn + t + f * msf] = 0.0f;
                        if (MASK == kMaskReim) gi[(long long)n * msn + t + f * msf] = 0.0f;
                    }
                }
                continue;
            }
            const float* wr = wave + (long long)n * wave_stride;
            const float* mr = mask_r ? mask_r + (long long)n * msn : nullptr;
            const float* mi = mask_i ? mask_i + (long long)n * msn : nullptr;
            if (!bwd)
                emu_tile<W, MASK, false>(wr, len, t0, window, vec_ok != 0, ft, mr, mi, (unsigned)msf, nullptr,
                                         out + nb, som, nullptr, nullptr, tmax, T, S);
            else
                emu_tile<W, MASK, true>(wr, len, t0, window, vec_ok != 0, bt, mr, mi, (unsigned)msf, dE + nb,
                                        nullptr, som, gr ? gr + (long long)n * msn : nullptr, gi ? gi + (long long)n * msn : nullptr,
                                        tmax, T, S, gwave ? gwave + (long long)n * wave_stride : nullptr);
        }
}

template <int W>
int emu_k1_w(int bwd, int mask_mode, const float* wave, const int* lengths, int n_utt,
             long long wave_stride, const float* mask_r, const float* mask_i,
             long long msn, long long msf, const float* window, const float* mel, int n_mels,
             float* out, const float* dE, float* gr, float* gi, int tmax, int vec_ok, float* gwave) {
    FwdTab ft;
    BwdTab bt;
    int ml[kBins];
    if (build_fwd_tab(mel, n_mels, &ft, ml) != 0) return -5;
    build_bwd_tab(ft, ml, &bt);
    set_warp_ranges(&ft, ml, W);
    switch (mask_mode) {
        case kMaskNone:  emu_k1_impl<W, kMaskNone>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, ft, bt, out, dE, gr, gi, tmax, vec_ok, gwave); break;
        case kMaskReim:  emu_k1_impl<W, kMaskReim>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, ft, bt, out, dE, gr, gi, tmax, vec_ok, gwave); break;
        case kMaskPower: emu_k1_impl<W, kMaskPower>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, ft, bt, out, dE, gr, gi, tmax, vec_ok, gwave); break;
        default: return -4;
    }
    return 0;
}

}  // namespace

// whole K1 (forward / backward), tile by tile, same control flow as lmfb_k1<>, for `warps` in {1,2,3,4,5}
extern "C" int emu_k1(int warps, int bwd, int mask_mode, const float* wave, const int* lengths, int n_utt,
                      long long wave_stride, const float* mask_r, const float* mask_i,
                      long long msn, long long msf, const float* window, const float* mel, int n_mels,
                      float* out, const float* dE, float* gr, float* gi, int tmax, int vec_ok, float* gwave) {
#define CALL(W) return emu_k1_w<W>(bwd, mask_mode, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mel, n_mels, out, dE, gr, gi, tmax, vec_ok, gwave)
    switch (warps) {
        case 1: CALL(1);
        case 2: CALL(2);
        case 3: CALL(3);
        case 4: CALL(4);
        case 5: CALL(5);
    }
#undef CALL
    return -3;
}

extern "C" int emu_mel_band(const float* mel, int n_mels, int warps, float* wl, float* wh, int* ml, int* lohi) {
    FwdTab ft;
    const int rc = build_fwd_tab(mel, n_mels, &ft, ml);
    if (rc != 0) return rc;
    set_warp_ranges(&ft, ml, warps);
    for (int f = 0; f < kBins; ++f) { wl[f] = ft.w[f].x; wh[f] = ft.w[f].y; }
    for (int w = 0; w < kMaxW; ++w) { lohi[2 * w] = ft.lo[w]; lohi[2 * w + 1] = ft.hi[w]; }
    int m = ml[0];                      // the advance counts must reproduce ml
    for (int f = 1; f < kBins; ++f) {
        m += (int)ft.adv[f];
        if (m != ml[f]) return -100 - f;
        if (((ft.hmask[f >> 3] >> (f & 7)) & 1) != (ft.adv[f] != 0)) return -300 - f;
    }
    return 0;
}

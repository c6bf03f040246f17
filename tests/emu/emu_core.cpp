// CPU emulation of the per-thread device code in aas_enhancement_b200/csrc/lmfb_core.cuh.
// TEST INFRASTRUCTURE ONLY: lets the index algebra of the lane=frame FFT be checked on a
// machine without a GPU.  A CTA of W warps is emulated phase by phase (a block barrier becomes
// "finish the phase for every warp and lane before starting the next one").
#include <vector>
#include <cmath>
#include <cstring>
#include "lmfb_core.cuh"
#include "mel_band.hpp"

using namespace aas_lmfb;

namespace {

template <int W, int MASK, bool BWD>
void emu_tile(const float* wave_row, int len, int t0, const float* window, bool vec_ok,
              const typename TabOf<BWD>::Param& tab, const float* mr, const float* mi, unsigned msf,
              const float* dE, float* out, unsigned som, float* gr, float* gi, int tmax, int T,
              std::vector<float2>& S, float* gwave_row = nullptr, const float* mel_dev = nullptr,
              bool first_ch = true, bool last_ch = true) {
    typename TabOf<BWD>::Smem sm;
    tables_fill(&sm, tab, 0, 1);
    window_fill(S.data(), window, 0, 1);
    std::vector<float> raw((kTile + 1) * kRawPitch, NAN);      // (uninitialised on the device)
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane) {
            const int n_rows = (T - t0 < kTile ? T - t0 : kTile) + 1;
            stage_raw<W, false>(w, lane, wave_row, len, t0, n_rows, raw.data(), nullptr, vec_ok);
        }
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane)
            fft_pass1<W>(w, raw.data() + lane * kRawPitch, S.data() + lane, S.data() + kTile);
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
                stage_lane_init(lane, sl);
                const int n_rows = (T - t0 < kTile ? T - t0 : kTile) + 1;
                unstage_tile<W>(w, lane, sl, gwave_row, len, t0, n_rows, Sin.data(), vec_ok);
            }
    }
    if constexpr (!BWD) {
        for (int w = 0; w < W; ++w)
            for (int lane = 0; lane < 32; ++lane) {
                const int t = t0 + lane;
                float* pl = reinterpret_cast<float*>(S.data()) + lane;
                if (tab.walkable) phase3_own<W>(w, pl, sm, tab, out + t, som * 4u, t < tmax, t < T, first_ch, last_ch);
                else phase3_gather<W>(w, pl, sm, tab.n_mels, mel_dev, out + t, som * 4u, t < tmax, t < T, first_ch, last_ch);
            }
    }
}

template <int W, int MASK>
void emu_k1_impl(int bwd, const float* wave, const int* lengths, int n_utt, long long wave_stride,
                 const float* mask_r, const float* mask_i, long long msn, long long msf,
                 const float* window, const FwdTab& ft, const BwdTab& bt, float* out, const float* dE,
                 float* gr, float* gi, int tmax, int vec_ok, float* gwave, const float* mel, int n_ch,
                 long long wave_stride_ch) {
    const int tiles = (tmax + kTile - 1) / kTile;
    const int n_mels = bwd ? bt.n_mels : ft.n_mels;
    std::vector<float2> S(kScratchBytes / 8);
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
                    else if (MASK != kMaskNone) for (int f = 0; f < kBins * n_ch; ++f) {
                        gr[(long long)n * msn + t + f * msf] = 0.0f;
                        if (MASK == kMaskReim) gi[(long long)n * msn + t + f * msf] = 0.0f;
                    }
                }
                continue;
            }
            for (int ch = 0; ch < n_ch; ++ch) {
                const float* wr = wave + (long long)n * wave_stride + (long long)ch * wave_stride_ch;
                const long long mo = (long long)n * msn + (long long)ch * kBins * msf;
                const float* mr = mask_r ? mask_r + mo : nullptr;
                const float* mi = mask_i ? mask_i + mo : nullptr;
                if (!bwd)
                    emu_tile<W, MASK, false>(wr, len, t0, window, vec_ok != 0, ft, mr, mi, (unsigned)msf, nullptr,
                                             out + nb, som, nullptr, nullptr, tmax, T, S, nullptr, mel, ch == 0, ch == n_ch - 1);
                else
                    emu_tile<W, MASK, true>(wr, len, t0, window, vec_ok != 0, bt, mr, mi, (unsigned)msf, dE + nb,
                                            nullptr, som, gr ? gr + mo : nullptr, gi ? gi + mo : nullptr,
                                            tmax, T, S, gwave ? gwave + (wr - wave) : nullptr);
            }
        }
}

template <int W>
int emu_k1_w(int bwd, int mask_mode, const float* wave, const int* lengths, int n_utt,
             long long wave_stride, const float* mask_r, const float* mask_i,
             long long msn, long long msf, const float* window, const float* mel, int n_mels,
             float* out, const float* dE, float* gr, float* gi, int tmax, int vec_ok, float* gwave,
             int n_ch, long long wave_stride_ch) {
    FwdTab ft;
    BwdTab bt;
    int ml[kBins];
    build_fwd_tab(mel, n_mels, &ft, ml);
    set_warp_ranges(&ft, ml, W);
    std::vector<float> dP;
    if (build_bwd_tab(mel, n_mels, &bt) != 0) {               // generic basis: dP = 1/4 B^T dE through the identity table
        build_bwd_tab_identity(&bt);
        if (bwd) {
            dP.assign((size_t)n_utt * kDpRows * tmax, 0.0f);
            for (int n = 0; n < n_utt; ++n)
                for (int f = 0; f < kBins; ++f)
                    for (int t = 0; t < tmax; ++t) {
                        float acc = 0.0f;
                        for (int m = 0; m < n_mels; ++m)
                            acc = fmaf(mel[m * kBins + f], dE[((size_t)n * n_mels + m) * tmax + t], acc);
                        dP[((size_t)n * kDpRows + f) * tmax + t] = 0.25f * acc;
                    }
            dE = dP.data();
        }
    }
    switch (mask_mode) {
        case kMaskNone:  emu_k1_impl<W, kMaskNone>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, ft, bt, out, dE, gr, gi, tmax, vec_ok, gwave, mel, n_ch, wave_stride_ch); break;
        case kMaskReim:  emu_k1_impl<W, kMaskReim>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, ft, bt, out, dE, gr, gi, tmax, vec_ok, gwave, mel, n_ch, wave_stride_ch); break;
        case kMaskPower: emu_k1_impl<W, kMaskPower>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, ft, bt, out, dE, gr, gi, tmax, vec_ok, gwave, mel, n_ch, wave_stride_ch); break;
        default: return -4;
    }
    return 0;
}

}  // namespace

// whole K1 (forward / backward), tile by tile, same control flow as lmfb_k1<>, for `warps` in {1,2,3,4,5}
extern "C" int emu_k1(int warps, int bwd, int mask_mode, const float* wave, const int* lengths, int n_utt,
                      long long wave_stride, const float* mask_r, const float* mask_i,
                      long long msn, long long msf, const float* window, const float* mel, int n_mels,
                      float* out, const float* dE, float* gr, float* gi, int tmax, int vec_ok, float* gwave,
                      int n_ch, long long wave_stride_ch) {
#define CALL(W) return emu_k1_w<W>(bwd, mask_mode, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mel, n_mels, out, dE, gr, gi, tmax, vec_ok, gwave, n_ch, wave_stride_ch)
    switch (warps) {
        case 1: CALL(1);
        case 2: CALL(2);
        case 3: CALL(3);
        case 4: CALL(4);
        case 5: CALL(5);
        case 6: CALL(6);
        case 8: CALL(8);
    }
#undef CALL
    return -3;
}

// the tables of a basis, rebuilt into dense matrices (what the kernels will effectively apply)
extern "C" int emu_mel_tables(const float* mel, int n_mels, int warps, float* fwd_dense, float* bwd_dense,
                              int* walkable, int* banded, int* lohi) {
    FwdTab ft;
    BwdTab bt;
    int ml[kBins];
    build_fwd_tab(mel, n_mels, &ft, ml);
    set_warp_ranges(&ft, ml, warps);
    *walkable = ft.walkable;
    for (int i = 0; i < n_mels * kBins; ++i) fwd_dense[i] = bwd_dense[i] = 0.0f;
    if (ft.walkable) {
        for (int f = 0; f < kBins; ++f) {
            if (ml[f] < n_mels) fwd_dense[ml[f] * kBins + f] += 4.0f * ft.w[f].x;
            if (ml[f] + 1 < n_mels) fwd_dense[(ml[f] + 1) * kBins + f] += 4.0f * ft.w[f].y;
        }
        int m = ml[0];                      // the advance counts must reproduce ml
        for (int f = 1; f < kBins; ++f) {
            if (ml[f] < ml[f - 1]) return -200 - f;
            m += (int)ft.adv[f];
            if (m != ml[f]) return -100 - f;
            if (((ft.hmask[f >> 3] >> (f & 7)) & 1) != (ft.adv[f] != 0)) return -300 - f;
        }
        for (int w = 0; w < kMaxW; ++w) { lohi[2 * w] = ft.lo[w]; lohi[2 * w + 1] = ft.hi[w]; }
    } else {
        for (int m = 0; m < n_mels; ++m) {
            const int lo = (int)(ft.row[m] & 255u), cnt = (int)(ft.row[m] >> 8);
            if (lo + cnt > kBins) return -10 - m;
            for (int i = 0; i < cnt; ++i) fwd_dense[m * kBins + lo + i] = mel[m * kBins + lo + i];
        }
    }
    *banded = build_bwd_tab(mel, n_mels, &bt) == 0 ? 1 : 0;
    if (*banded)
        for (int k2 = 0; k2 < 17; ++k2)
            for (int k1 = 0; k1 < 5; ++k1)
                for (int side = 0; side < 2; ++side) {
                    const int f0 = bin_of(k2, k1), f = side ? kBins - 1 - f0 : f0;
                    const int d = (int)bt.d[k2][k1][side];
                    if (d < 0 || d + 1 >= n_mels) return -400 - f;
                    bwd_dense[d * kBins + f] = 4.0f * bt.w[k2][k1][2 * side];
                    bwd_dense[(d + 1) * kBins + f] = 4.0f * bt.w[k2][k1][2 * side + 1];
                }
    return 0;
}

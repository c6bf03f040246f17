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
              const MelBand& mb, const float* mr, const float* mi, unsigned /*sf*/,
              const float* dE, float* out, unsigned som, float* gr, float* gi, int tmax, int T,
              std::vector<float2>& S) {
    Tables tb;
    tables_fill(&tb, mb, mb.ent[1].moff, 0, 1);            // ent[1].moff == bytes per mask row
    window_fill(S.data(), window, 0, 1);
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane) {
            StageLane sl;
            stage_lane_init(lane, sl);
            const int n_rows = (T - t0 < kTile ? T - t0 : kTile) + 1;
            stage_tile<W>(w, lane, sl, wave_row, len, t0, n_rows, S.data(), vec_ok);
        }
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane) fft_pass1<W>(w, S.data() + lane, S.data() + kTile);
    for (int w = 0; w < W; ++w)
        for (int lane = 0; lane < 32; ++lane) {
            const int t = t0 + lane;
            const bool inrow = t < tmax;
            const long long clamp = inrow ? 0 : (long long)(tmax - 1 - t);
            const float* mrp = mr ? mr + t + clamp : nullptr;
            const float* mip = mi ? mi + t + clamp : nullptr;
            const float* dep = dE ? dE + t + clamp : nullptr;
            StepMasks first;
            load_masks<MASK, BWD>(w, tb, mrp, mip, first);
            fft_pass2<W, MASK, BWD, false>(w, S.data() + lane, tb, first, mrp, mip, dep, som * 4u,
                                    gr ? gr + t : nullptr, gi ? gi + t : nullptr, inrow);
        }
    if (!BWD)
        for (int w = 0; w < W; ++w)
            for (int lane = 0; lane < 32; ++lane) {
                const int t = t0 + lane;
                phase3_fwd(w, S.data() + lane, tb, out + t, som * 4u, t < tmax, t < T);
            }
}

template <int W, int MASK>
void emu_k1_impl(int bwd, const float* wave, const int* lengths, int n_utt, long long wave_stride,
                 const float* mask_r, const float* mask_i, long long msn, long long msf,
                 const float* window, const MelBand& mb, float* out, const float* dE,
                 float* gr, float* gi, int tmax, int vec_ok) {
    const int tiles = (tmax + kTile - 1) / kTile;
    const int n_mels = mb.n_mels;
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
                emu_tile<W, MASK, false>(wr, len, t0, window, vec_ok != 0, mb, mr, mi, (unsigned)msf, nullptr,
                                         out + nb, som, nullptr, nullptr, tmax, T, S);
            else
                emu_tile<W, MASK, true>(wr, len, t0, window, vec_ok != 0, mb, mr, mi, (unsigned)msf, dE + nb,
                                        nullptr, som, gr + (long long)n * msn, gi ? gi + (long long)n * msn : nullptr,
                                        tmax, T, S);
        }
}

template <int W>
int emu_k1_w(int bwd, int mask_mode, const float* wave, const int* lengths, int n_utt,
             long long wave_stride, const float* mask_r, const float* mask_i,
             long long msn, long long msf, const float* window, const float* mel, int n_mels,
             float* out, const float* dE, float* gr, float* gi, int tmax, int vec_ok) {
    MelBand mb, mbb;
    memset(&mb, 0, sizeof(mb));
    int ml[kBins];
    uint8_t dlo[kBins];
    if (build_mel_band(mel, n_mels, W, &mb, ml) != 0) return -5;
    if (bwd) {
        make_bwd_band(mb, ml, &mbb, dlo);
        mb = mbb;
        patch_strides(&mb, (unsigned)msf, dlo, (unsigned)tmax);
    } else {
        patch_strides(&mb, (unsigned)msf, nullptr, 0);
    }
    switch (mask_mode) {
        case kMaskNone:  emu_k1_impl<W, kMaskNone>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mb, out, dE, gr, gi, tmax, vec_ok); break;
        case kMaskReim:  emu_k1_impl<W, kMaskReim>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mb, out, dE, gr, gi, tmax, vec_ok); break;
        case kMaskPower: emu_k1_impl<W, kMaskPower>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mb, out, dE, gr, gi, tmax, vec_ok); break;
        default: return -4;
    }
    return 0;
}

}  // namespace

// whole K1 (forward / backward), tile by tile, same control flow as lmfb_k1<>, for `warps` in {1,2,4,5}
extern "C" int emu_k1(int warps, int bwd, int mask_mode, const float* wave, const int* lengths, int n_utt,
                      long long wave_stride, const float* mask_r, const float* mask_i,
                      long long msn, long long msf, const float* window, const float* mel, int n_mels,
                      float* out, const float* dE, float* gr, float* gi, int tmax, int vec_ok) {
#define CALL(W) return emu_k1_w<W>(bwd, mask_mode, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mel, n_mels, out, dE, gr, gi, tmax, vec_ok)
    switch (warps) {
        case 1: CALL(1);
        case 2: CALL(2);
        case 4: CALL(4);
        case 5: CALL(5);
    }
#undef CALL
    return -3;
}

extern "C" int emu_mel_band(const float* mel, int n_mels, int warps, float* wl, float* wh, int* ml, int* mbeg) {
    MelBand mb;
    memset(&mb, 0, sizeof(mb));
    const int rc = build_mel_band(mel, n_mels, warps, &mb, ml);
    for (int f = 0; f < kBins; ++f) { wl[f] = mb.ent[f].wl; wh[f] = mb.ent[f].wh; }
    for (int w = 0; w <= kMaxW; ++w) mbeg[w] = mb.mbeg[w];
    if (rc == 0) {                      // the filter ranges must tile the bins in order
        int f = 0;
        for (int m = 0; m < n_mels; ++m)
            for (; f < mb.fend[m]; ++f) if (ml[f] != m) return -100 - m;
        for (; f < kBins; ++f) if (ml[f] != n_mels) return -300;
        for (int w = 0; w < warps; ++w) if (mb.mbeg[w] > mb.mbeg[w + 1]) return -400;
        if (mb.mbeg[0] != 0 || mb.mbeg[warps] != n_mels) return -401;
    }
    return rc;
}

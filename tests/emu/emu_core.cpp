// CPU emulation of the per-thread device code in aas_enhancement_b200/csrc/lmfb_core.cuh.
// TEST INFRASTRUCTURE ONLY: lets the index algebra of the lane=frame FFT be checked on a
// machine without a GPU.  A warp is emulated by running the 32 lanes one after another.
#include <vector>
#include <cstring>
#include "lmfb_core.cuh"

using namespace aas_lmfb;

extern "C" int emu_stft_tile(const float* wave_row, int len, int t0, const float* window,
                             int vec_ok, float* re_out /*[161][32]*/, float* im_out /*[161][32]*/) {
    std::vector<float2> S(kSlots * kPitch);
    // power-spectrum probes: 'reim' forward with masks (1,0) and (0,1) gives Re'^2 and Im'^2
    std::vector<float> ones(kBins, 1.0f), zeros(kBins, 0.0f);
    for (int pass = 0; pass < 2; ++pass) {
        for (int lane = 0; lane < 32; ++lane) {
            StageLane sl;
            stage_lane_init(lane, window, sl);
            stage_tile(lane, sl, wave_row, len, t0, S.data(), vec_ok != 0);
        }
        for (int lane = 0; lane < 32; ++lane) {
            float2* col = S.data() + lane;
            fft_pass1(col);
            fft_pass2_masked<kMaskReim, false>(col, pass == 0 ? ones.data() : zeros.data(),
                                               pass == 0 ? zeros.data() : ones.data(), 1u);
            const float* colf = reinterpret_cast<const float*>(col);
            for (int f = 0; f < kBins; ++f) (pass == 0 ? re_out : im_out)[f * 32 + lane] = colf[kBinOff[f]];
        }
    }
    return 0;
}

// ---- whole K1 (forward / backward), tile by tile, same control flow as lmfb_k1<> ----
#include "mel_band.hpp"

template <int MASK>
static void emu_k1_impl(int bwd, const float* wave, const int* lengths, int n_utt, long long wave_stride,
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
            const long long som = tmax;
            if (t0 >= T) {
                for (int lane = 0; lane < 32; ++lane) {
                    const int t = t0 + lane;
                    if (t >= tmax) continue;
                    const long long row_nm = (long long)n * n_mels * som + t;
                    if (!bwd) for (int m = 0; m < n_mels; ++m) out[row_nm + m * som] = 0.0f;
                    else if (MASK != kMaskNone) for (int f = 0; f < kBins; ++f) {
                        gr[(long long)n * msn + t + f * msf] = 0.0f;
                        if (MASK == kMaskReim) gi[(long long)n * msn + t + f * msf] = 0.0f;
                    }
                }
                continue;
            }
            for (int lane = 0; lane < 32; ++lane) {
                StageLane sl;
                stage_lane_init(lane, window, sl);
                stage_tile(lane, sl, wave + (long long)n * wave_stride, len, t0, S.data(), vec_ok != 0);
            }
            for (int lane = 0; lane < 32; ++lane) {
                const int t = t0 + lane;
                const bool inrow = t < tmax, valid = t < T;
                const long long row_nm = (long long)n * n_mels * som + t;
                const long long moff = (long long)n * msn + t;
                const long long clamp = inrow ? 0 : (long long)(tmax - 1 - t);
                const bool has_mask = MASK != kMaskNone;
                float2* col = S.data() + lane;
                fft_pass1(col);
                if (!bwd) {
                    fft_pass2_masked<MASK, false>(col, mask_r + (has_mask ? moff + clamp : 0),
                                                  mask_i + (MASK == kMaskReim ? moff + clamp : 0), (unsigned)msf);
                    phase3_fwd(col, mb, out + row_nm, (unsigned)som, inrow, valid);
                } else {
                    float dw[kDWin];
                    dwin_preload(dE + row_nm + clamp, (unsigned)som, n_mels, dw);
                    fft_pass2_masked<MASK, true>(col, mask_r + (has_mask ? moff + clamp : 0),
                                                 mask_i + (MASK == kMaskReim ? moff + clamp : 0), (unsigned)msf);
                    phase3_bwd<MASK>(col, mb, dE + row_nm + clamp, (unsigned)som, dw, gr + moff, gi + moff,
                                     (unsigned)msf, inrow);
                }
            }
        }
}

extern "C" int emu_k1(int bwd, int mask_mode, const float* wave, const int* lengths, int n_utt,
                      long long wave_stride, const float* mask_r, const float* mask_i,
                      long long msn, long long msf, const float* window, const float* mel, int n_mels,
                      float* out, const float* dE, float* gr, float* gi, int tmax, int vec_ok) {
    MelBand mb;
    memset(&mb, 0, sizeof(mb));
    if (build_mel_band(mel, n_mels, &mb) != 0) return -5;
    static const float zero = 0.0f;
    if (!mask_r) mask_r = &zero;      // never dereferenced in the modes that leave it NULL
    if (!mask_i) mask_i = &zero;
    switch (mask_mode) {
        case kMaskNone:  emu_k1_impl<kMaskNone>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mb, out, dE, gr, gi, tmax, vec_ok); break;
        case kMaskReim:  emu_k1_impl<kMaskReim>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mb, out, dE, gr, gi, tmax, vec_ok); break;
        case kMaskPower: emu_k1_impl<kMaskPower>(bwd, wave, lengths, n_utt, wave_stride, mask_r, mask_i, msn, msf, window, mb, out, dE, gr, gi, tmax, vec_ok); break;
        default: return -4;
    }
    return 0;
}

extern "C" int emu_mel_band(const float* mel, int n_mels, float* wl, float* wh, int* ml) {
    MelBand mb;
    memset(&mb, 0, sizeof(mb));
    const int rc = build_mel_band(mel, n_mels, &mb, ml);
    for (int f = 0; f < kBins; ++f) { wl[f] = mb.ent[f].wl; wh[f] = mb.ent[f].wh; }
    // the filter ranges must tile the bins in order
    if (rc == 0) {
        int f = 0;
        for (int m = 0; m < n_mels; ++m) {
            for (; f < mb.fend[m]; ++f) if (ml[f] != m) return -100 - m;
        }
        for (; f < kBins; ++f) if (ml[f] != n_mels) return -300;
    }
    return rc;
}

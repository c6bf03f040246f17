/* aas_lmfb.h -- C ABI of the B200-native LMFB front-end (libaas_lmfb.so).
 *
 * The reference (lifelongeek/AAS_enhancement) has no FFI layer; its de-facto operator API for
 * this path is Python:
 *   - BRNNmultiCH.__init__(..., mel_basis) / .forward   Speech_enhancement_by_AAS/model.py:148-200
 *     (mask -> power -> mel -> log1p glue at :186-198, mel buffer at :167)
 *   - the front-end parameters                          AM_training/train.py:39-42, :55-61, :190-199
 *   - the (N, 40, T) fp32 channel-first batches         Speech_enhancement_by_AAS/loader_functions.py:47-105
 * Each entry point below names the reference code it stands in for.  Plain pointers and
 * sizes only; no torch types.  All device buffers are allocated and owned by the caller; the
 * library never allocates device memory, never synchronises, and never keeps a pointer past
 * the call.  Everything is asynchronous on the caller's stream.
 *
 * Layouts (fp32 unless noted):
 *   wave      (N, wave_stride)   zero-padded samples, 16 kHz mono
 *   lengths   (N,) int32         samples per utterance
 *   mask_r/i  (N, 161, Tmax)     element (n,f,t) at n*mask_stride_n + f*mask_stride_f + t
 *   out       (N, M, Tmax)       contiguous; frames t >= T_n are written as exact zeros
 *   stats     (N, M, 2)          (mean, rstd) written by forward, read by backward
 *   frame counts                 T_n = 1 + lengths[n] / 160, clipped to Tmax
 */
#ifndef AAS_LMFB_H
#define AAS_LMFB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AAS_LMFB_ABI_VERSION 2

#define AAS_LMFB_N_FFT   320   /* int(16000 * 0.02), AM_training/train.py:39-40 */
#define AAS_LMFB_HOP     160   /* int(16000 * 0.01), AM_training/train.py:41    */
#define AAS_LMFB_N_BINS  161   /* AM_training/train.py:199                      */

/* flags: bits 0-1 mask mode, bits 2-3 CMVN mode, bit 4 sample format */
#define AAS_LMFB_MASK_NONE    0u   /* P = Re^2 + Im^2                                        */
#define AAS_LMFB_MASK_REIM    1u   /* P = (Re*Mr)^2 + (Im*Mi)^2   (model.py:191-194)          */
#define AAS_LMFB_MASK_POWER   2u   /* P = M * (Re^2 + Im^2)                                   */
#define AAS_LMFB_CMVN_NONE    (0u << 2)
#define AAS_LMFB_CMVN_PER_BIN (1u << 2)   /* mean/std per mel bin over the utterance's frames */
#define AAS_LMFB_CMVN_GLOBAL  (2u << 2)   /* scalar mean/std over the utterance's (M, T) matrix */
#define AAS_LMFB_WAVE_I16     (1u << 4)   /* `wave` is int16 PCM (value / 32768): what wave files hold
                                           * (loader.py save_wave(int16=True)); halves the bytes on the wire
                                           * and in HBM; converted inside the kernel; no waveform gradient   */

/* return codes: 0 ok, <0 argument errors, >0 a cudaError_t from the launch */
#define AAS_LMFB_OK              0
#define AAS_LMFB_E_NULL         -1
#define AAS_LMFB_E_ALIGN        -2
#define AAS_LMFB_E_SHAPE        -3
#define AAS_LMFB_E_FLAGS        -4
#define AAS_LMFB_E_MEL          -5   /* generic mel basis: the call needs io.mel_dev (see plan_create) */
#define AAS_LMFB_E_NOMEM        -6   /* host allocation failed                                 */

typedef struct aas_lmfb_plan aas_lmfb_plan;

int aas_lmfb_abi_version(void);
const char* aas_lmfb_strerror(int code);

/* Build a host-side plan from the mel basis, a HOST (n_mels, 161) row-major fp32 matrix -- the
 * `mel_basis` constructor argument of BRNNmultiCH (model.py:148, :167).  ANY matrix is accepted, as
 * the reference's k=1 conv1d accepts any (model.py:196).  Triangular filterbanks (every bin feeds at
 * most two adjacent filters, rows short enough for the on-chip table) run on the fast path; other
 * bases (dense, re-ordered, hand-made) run on a generic path that needs the caller's DEVICE copy of
 * the same matrix (`mel_dev` of aas_lmfb_io) and more workspace.  No device work here. */
aas_lmfb_plan* aas_lmfb_plan_create(const float* mel_host, int n_mels, int n_bins, int* status);
void aas_lmfb_plan_destroy(aas_lmfb_plan* plan);
/* What the analysis found: *fwd_compact / *bwd_banded are 1 on the fast paths (any may be NULL). */
int aas_lmfb_plan_info(const aas_lmfb_plan* plan, int* n_mels, int* fwd_compact, int* bwd_banded);
/* Tuning knobs (tests and benchmarks; 0 = the measured default): warps per 32-frame tile of the
 * forward / backward kernel (4..6), and static_schedule = 1 to deal tiles round-robin instead of
 * handing them out with cluster launch control.  Results do not depend on them. */
int aas_lmfb_plan_set_tuning(aas_lmfb_plan* plan, int warps_fwd, int warps_bwd, int static_schedule);

/* Optional, recommended: a device-resident copy of the plan's lookup tables.  The caller allocates
 * aas_lmfb_plan_tables_bytes() bytes (16-byte aligned) on the device it will run on, has them filled
 * ONCE with aas_lmfb_plan_upload (an asynchronous host-to-device copy on `cuda_stream`; the plan must
 * outlive it) and passes the pointer as aas_lmfb_io.tables on every call.  Without it every thread
 * block rebuilds the tables from kernel parameters (about 4 us per block: a quarter of a small launch). */
size_t aas_lmfb_plan_tables_bytes(const aas_lmfb_plan* plan);
int aas_lmfb_plan_upload(const aas_lmfb_plan* plan, void* tables_dev, void* cuda_stream);

/* Bytes of device workspace `backward` needs for this plan (forward needs none). */
size_t aas_lmfb_workspace_bytes(const aas_lmfb_plan* plan, int n, int tmax, uint32_t flags);

/* ---- extended call: everything in one struct -------------------------------------------------
 * Multi-channel input (BRNNmultiCH with nCH > 1, model.py:160-167, :186-198: masks are
 * (N, nCH*161, T), the basis repeats over the channels, i.e. the masked powers of the channels are
 * summed before the mel projection), generic bases, an explicit device and the readable length of
 * the wave rows.  The classic entry points below are thin wrappers around these. */
typedef struct aas_lmfb_io {
    uint32_t       struct_size;     /* sizeof(aas_lmfb_io): lets the struct grow compatibly            */
    uint32_t       flags;           /* mask mode | CMVN mode                                           */
    int32_t        device;          /* CUDA device of the buffers; -1 = the calling thread's current   */
    int32_t        n;               /* utterances                                                      */
    int32_t        n_ch;            /* channels per utterance, >= 1                                    */
    int32_t        tmax;
    float          eps;
    int32_t        reserved_;
    const void*    wave;            /* fp32 (or int16 with AAS_LMFB_WAVE_I16), (N, n_ch, .): sample (n, c, i) at
                                     * n*wave_stride + c*wave_stride_ch + i (strides in samples)         */
    int64_t        wave_stride;
    int64_t        wave_stride_ch;
    int64_t        wave_len;        /* samples of every row that may be read; lengths are clamped to it (0: trust lengths) */
    const int32_t* lengths;         /* (N,) samples; all channels of an utterance share it             */
    const float*   mask_r;          /* (N, n_ch*161, Tmax): row c*161 + f                              */
    const float*   mask_i;
    int64_t        mask_stride_n;
    int64_t        mask_stride_f;
    const float*   window;          /* device (320,)                                                   */
    const float*   mel_dev;         /* device (n_mels, 161), contiguous: needed only by generic plans  */
    float*         out;             /* (N, M, Tmax)                                                    */
    float*         stats;           /* (N, M, 2)                                                       */
    const float*   grad_out;        /* backward: (N, M, Tmax)                                          */
    float*         grad_mask_r;     /* backward: strides of the masks                                  */
    float*         grad_mask_i;
    float*         grad_wave;       /* backward, optional: layout of `wave`                            */
    void*          workspace;       /* backward: aas_lmfb_workspace_bytes() bytes, 16-byte aligned     */
    void*          cuda_stream;
    void* const*   prof;            /* optional 4 cudaEvent_t recorded around the two kernels          */
    const void*    tables;          /* optional: device tables filled by aas_lmfb_plan_upload          */
    const float*   l1_target;       /* forward, optional (needs a CMVN mode): (N, M, Tmax) target of the
                                     * L1Loss_mask that follows (model.py:19-31) ...                     */
    float*         l1_rows;         /* ... and (N, M) floats that receive sum_t |Z - target| of every row,
                                     * formed while Z is in registers (Z is read once); aas_l1_rows_sum
                                     * adds them up in a fixed order                                     */
    int32_t*       frame_lens;      /* forward, optional: (N,) receives the frame counts T_n                 */
} aas_lmfb_io;

int aas_lmfb_forward_ex(const aas_lmfb_plan* plan, const aas_lmfb_io* io);
int aas_lmfb_backward_ex(const aas_lmfb_plan* plan, const aas_lmfb_io* io);

/* Forward: framing + Hamming window + STFT(320/160) + mask + mel + log1p (+ CMVN).
 * Stands in for the missing SpectrogramDataset front-end (AM_training/train.py:190-199,
 * :255-259) fused with the glue of model.py:186-198.
 *   window   device (320,) window samples (AM_training/train.py:42)
 *   mask_r   may be NULL for MASK_NONE; mask_i may be NULL unless MASK_REIM
 *   prof     optional array of 4 cudaEvent_t recorded around {stft-mel kernel, cmvn kernel};
 *            NULL in normal use */
int aas_lmfb_forward(const aas_lmfb_plan* plan,
                     const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                     const float* mask_r, const float* mask_i,
                     int64_t mask_stride_n, int64_t mask_stride_f,
                     const float* window,
                     float* out, float* stats, int tmax,
                     uint32_t flags, float eps, void* cuda_stream, void* const* prof);

/* Backward into the mask(s): what autograd does through model.py:191-198 (plus the CMVN).
 *   out, stats   exactly what forward produced (not modified; backward may run repeatedly,
 *                trainer_AAS.py:150 uses retain_graph=True)
 *   grad_out     (N, M, Tmax) contiguous
 *   grad_mask_r/i same strides as the masks; every element (incl. frames t >= T_n) is written
 *   workspace    aas_lmfb_workspace_bytes() bytes, 16-byte aligned */
int aas_lmfb_backward(const aas_lmfb_plan* plan,
                      const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                      const float* mask_r, const float* mask_i,
                      int64_t mask_stride_n, int64_t mask_stride_f,
                      const float* window,
                      const float* out, const float* stats, const float* grad_out,
                      float* grad_mask_r, float* grad_mask_i,
                      void* workspace, int tmax,
                      uint32_t flags, float eps, void* cuda_stream, void* const* prof);

/* Backward that ALSO returns the gradient w.r.t. the waveform (second half):
 * what autograd would give a waveform-domain enhancer feeding this front-end.  Same arguments as
 * aas_lmfb_backward plus
 *   grad_wave  (N, wave_stride) fp32, 8-byte aligned, same layout as `wave`; the library zeroes
 *              the first min(wave_stride, 160*tmax) samples of every row and accumulates into them
 *              (frames overlap, and the reflect padding folds the ends of an utterance back);
 *              samples past lengths[n] stay zero.
 * With MASK_NONE the mask pointers and grad_mask_r/i may be NULL (only the waveform takes a
 * gradient).  The result is deterministic up to the order of the few additions per sample. */
int aas_lmfb_backward_wave(const aas_lmfb_plan* plan,
                           const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                           const float* mask_r, const float* mask_i,
                           int64_t mask_stride_n, int64_t mask_stride_f,
                           const float* window,
                           const float* out, const float* stats, const float* grad_out,
                           float* grad_mask_r, float* grad_mask_i, float* grad_wave,
                           void* workspace, int tmax,
                           uint32_t flags, float eps, void* cuda_stream);

/* STFT as an output (first half): the input BRNNmultiCH.forward takes,
 * `(N, nCH*F*2, T)` with the F real rows first and the F imaginary rows second
 * (Speech_enhancement_by_AAS/model.py:170, :186-188), computed from the waveform with the same
 * framing / window / 320-point FFT as aas_lmfb_forward (AM_training/train.py:39-42, :190-199).
 *   out  (N, 2, 161, Tmax) fp32, element (n, c, f, t) at n*out_stride_n + (c*161 + f)*tmax + t;
 *        frames t >= T_n are written as exact zeros.  No mel basis is involved (any plan will do). */
int aas_lmfb_stft(const aas_lmfb_plan* plan,
                  const float* wave, const int32_t* lengths, int n, int64_t wave_stride,
                  const float* window, float* out, int64_t out_stride_n, int tmax,
                  void* cuda_stream);

/* ---- L1Loss_mask: the loss applied to the features right after this front-end --------------
 * Replaces Speech_enhancement_by_AAS/model.py:19-31 (called at trainer_AAS.py:146-161, :176-181,
 * trainer_DCE.py, trainer_FSEGAN.py): sum |a - b| over (N, C, Tmax), deterministic two-stage sum.
 *   mask  NULL reproduces the reference exactly (its masked_fill is a no-op, so padded frames
 *         contribute); a (N, 1, Tmax) byte mask (1 = padding) zeroes the padded frames instead.
 *   partial  scratch of aas_l1_partial_count() floats;  out  one float (the un-normalised sum).
 * The division by nElement = numel(mask) - sum(mask) (frames, not elements) is host-side glue. */
int aas_l1_partial_count(void);
int aas_l1_abs_sum(const float* a, const float* b, const uint8_t* mask, int n, int c, int tmax,
                   float* partial, float* out, void* cuda_stream);
/* Sum of `count` per-row partial sums (aas_lmfb_io.l1_rows) into one float, deterministic. */
int aas_l1_rows_sum(const float* rows, int count, float* out, void* cuda_stream);
/* d/da and d/db of scale * sum|a - b|: grad_a = scale * sign(a - b), grad_b = -grad_a (either may be
 * NULL); `scale` is a device scalar so that no host sync is needed. */
int aas_l1_abs_grad(const float* a, const float* b, const uint8_t* mask, int n, int c, int tmax,
                    const float* scale, float* grad_a, float* grad_b, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* AAS_LMFB_H */

/*
 * frt.h -- C ABI of friture_b200, the B200-native (sm_100a) implementation of Friture's
 * per-chunk spectral hot path.
 *
 * The reference (tlecomte/friture) is pure Python/NumPy and has no FFI: its boundary is the
 * Python call surface the widgets use.  Each entry point below names the reference interface
 * it stands behind (file:line relative to the reference root).  The host side
 * (friture_b200/*.py) binds these with ctypes and keeps the reference's class / method names.
 *
 * Conventions
 *   - plain pointers and sizes only; `*_dev` pointers are device memory on the handle's GPU,
 *     `*_host` pointers are host memory (pinned for full PCIe speed, pageable also works);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); device
 *     entry points are asynchronous on it, host entry points return when the result is in
 *     the caller's host buffer;
 *   - the library never owns caller buffers; plans, coefficient tables and filter / smoothing
 *     state live in the handle;
 *   - every function returns FRT_OK (0) or a negative FRT_E* code; frt_last_error(h) gives the
 *     message.  There is no CPU fallback anywhere: without a CUDA device frt_create fails.
 */
#ifndef FRT_H
#define FRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRT_OK 0
#define FRT_EINVAL (-1)   /* bad argument (shape, alignment, unsupported size)           */
#define FRT_ECUDA (-2)    /* CUDA runtime error (message holds cudaGetErrorString)       */
#define FRT_ENOMEM (-3)   /* host or device allocation failed                            */
#define FRT_ESTATE (-4)   /* call order problem (e.g. process before plan)               */

typedef struct frt_ctx *frt_handle;

/* ---------------------------------------------------------------- lifecycle */
int frt_version(void);
/* Bind a handle to CUDA device `device`.  Fails (FRT_ECUDA) when no device is usable.       */
int frt_create(int device, frt_handle *out);
int frt_destroy(frt_handle h);
const char *frt_last_error(frt_handle h);   /* h may be NULL: last frt_create failure      */
int frt_device_sm_count(frt_handle h);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches).            */
int64_t frt_launch_count(frt_handle h);
/* Pinned host memory for the *_host entry points.                                          */
int frt_host_alloc(frt_handle h, size_t bytes, void **out);
int frt_host_free(frt_handle h, void *p);

/* ---------------------------------------------------------------- STFT (audioproc)
 * Stands behind friture/audioproc.py:27-81 (`audioproc.set_fftsize`, `analyzelive`,
 * `norm_square`, `update_window`) and the framing loops that call it,
 * friture/spectrogram.py:131-159 and friture/spectrum.py:125-155, plus
 * `log_spectrogram`, friture/spectrogram.py:119-125.                                        */

#define FRT_STFT_POWER 0      /* |rfft(x*hann)|^2 / N^2           (analyzelive)             */
#define FRT_STFT_LOGPOWER 1   /* 10*log10(power + 1e-30)          (log_spectrogram)         */

/* Build the symmetric-Hann window and twiddle tables for `n_fft` (32*2^k, 32..16384;
 * friture/spectrum_settings.py:61-70).  Replaces audioproc.set_fftsize.                    */
int frt_stft_plan(frt_handle h, int n_fft);
/* Copy the plan's float32 window (n_fft values) to `window_host` (audioproc.window).        */
int frt_stft_window(frt_handle h, float *window_host);
/* Batched STFT.  Channel c, frame f is samples x[c*x_stride + f*hop .. + n_fft) and is written
 * to out[c*out_stride_c + f*out_stride_f + k], k = 0..n_fft/2.  One frame == one
 * analyzelive() call of the reference.                                                      */
int frt_stft_process(frt_handle h, const float *x_dev, int64_t x_stride, int n_channels,
                     int64_t n_frames, int hop, float *out_dev, int64_t out_stride_c,
                     int64_t out_stride_f, int mode, void *stream);
/* Same with host buffers: H2D copy, kernels and D2H copy are pipelined over channel groups
 * inside the call (bench.py's `e2e`).  n_samples is per channel; frames = (n_samples-n_fft)/hop+1,
 * out_host is [n_channels][frames][n_fft/2+1] contiguous.                                   */
int frt_stft_process_host(frt_handle h, const float *x_host, int64_t x_stride, int n_channels,
                          int64_t n_samples, int hop, float *out_host, int mode);

#ifdef __cplusplus
}
#endif
#endif /* FRT_H */

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

/* ---------------------------------------------------------------- filterbank (Octave_Filters)
 * Stands behind friture/octavefilters.py:37-158 (`Octave_Filters.filter`, `setbandsperoctave`)
 * with the numerics of the reference's IIR bank `octave_filter_bank_decimation`
 * (friture/filter.py:86-118, `decimate` friture/signal/decimate.py:27-42, recursion
 * friture/signal/lfilter.py:85-147), fused with the octave widget's smoothing
 * `exp_smoothed_value(kernel, alpha, y**2, old)` (friture/octavespectrum.py:104,140-156;
 * friture/signal/exp_smoothing.py:11-56) and `10*log10(sp+1e-30)` (octavespectrum.py:119-120). */

/* Build the filterbank for `n_channels` independent streams.
 *   sos_band  [bands_per_octave][2][6]  second-order sections (b0 b1 b2 1 a1 a2) of the
 *             band-passes of the top octave, band 0 = lowest (generated_filters.PARAMS[bpo])
 *   sos_dec   [6][6]                    sections of the decimation low-pass (PARAMS['dec'])
 *   alphas    [n_octaves]               smoothing factor per stage j (rate fs/2^j),
 *             octavespectrum.py:150-154
 * Filter and smoothing state start at zero (filter.py:121-133; octavespectrum.py:59).        */
int frt_bank_plan(frt_handle h, int n_channels, int bands_per_octave, int n_octaves,
                  const double *sos_band, const double *sos_dec, const double *alphas);
int frt_bank_reset(frt_handle h);
/* Per-band dB offsets added to the energies when db != 0 -- the octave widget's weighting,
 * `10*log10(sp+1e-30) + w` with w = Octave_Filters.A / .B / .C (friture/octavespectrum.py:108-121,
 * friture/octavefilters.py:76-82).  weight_db_host: nbands floats (band k as below), or NULL for
 * no weighting.  Belongs to the current plan (a new frt_bank_plan clears it).                   */
int frt_bank_set_weighting(frt_handle h, const float *weight_db_host);
/* The software-pipeline schedule of the fused-energy kernel for a stream of n_samples per channel
 * processed in steps of 2^log2_chunk (5 or 6) samples: stage_start[10] = step at which stage j
 * starts, *n_steps = steps until the pipeline has drained.  Pure host arithmetic (needs no GPU);
 * exported so that tests/bank_pipeline_model.py can check the model against the library.       */
int frt_bank_schedule(int n_octaves, int log2_chunk, int64_t n_samples, int *stage_start,
                      int *n_steps);
/* The same for either lane layout of the kernel: sections_per_lane = 2 (a half-warp per channel, what
 * frt_bank_schedule reports) or 1 (a warp per channel: band chains of two lanes, decimator chains of
 * six -- half the serial work per lane, the layout used for few channels).                         */
int frt_bank_schedule2(int n_octaves, int log2_chunk, int sections_per_lane, int64_t n_samples,
                       int *stage_start, int *n_steps);
/* Process n_blocks consecutive blocks of `block` samples per channel
 * (x[c*x_stride + b*block + n]); block % 256 == 0 (the reference needs even lengths at every
 * stage, decimate.py:41).  State is carried across calls, so any blocking of a stream gives
 * the same result.
 *   energies_dev  [C][n_blocks][nbands] or NULL: smoothed band energies after each block
 *                 (band k = (n_octaves-1-j)*bpo + i as in filter.py:104-109); dB when db != 0
 *   y_dev         NULL, or the band outputs themselves (the literal `.filter()` contract):
 *                 channel c at y_dev + c*y_stride, bands concatenated k = 0..nbands-1, band k
 *                 holding (block*n_blocks) >> j samples                                       */
int frt_bank_process(frt_handle h, const float *x_dev, int64_t x_stride, int block, int n_blocks,
                     float *energies_dev, float *y_dev, int64_t y_stride, int db, void *stream);
/* Same with the energies of channel c, block b written to energies_dev + c*e_stride_c + b*nbands
 * (e_stride_c >= n_blocks*nbands): lets a caller that cuts one stream into several launches fill
 * one [C][all blocks][nbands] array.  No ragged band outputs.                                  */
int frt_bank_process_strided(frt_handle h, const float *x_dev, int64_t x_stride, int block,
                             int n_blocks, float *energies_dev, int64_t e_stride_c, int db,
                             void *stream);
/* State checkpoint / resume (the reference keeps its state inside the object,
 * octavefilters.py:50-56).  z: [C][n_octaves][2*bpo+6][2] section states; ema:
 * [C][n_octaves][bpo] smoothed energy divided by alpha_j (the kernel's internal form, so that a
 * get/set round trip is bit-exact).  The sections run in normalised form (numerator 1, c, 1; the
 * chain gain is applied to the chain's output), z is the state of those sections.             */
int frt_bank_state_size(frt_handle h, int64_t *z_floats, int64_t *ema_floats);
int frt_bank_get_state(frt_handle h, float *z_host, float *ema_host);
int frt_bank_set_state(frt_handle h, const float *z_host, const float *ema_host);

/* ---------------------------------------------------------------- live-path filterbank (FIR)
 * The numerics of the reference's LIVE `Octave_Filters.filter` (friture/octavefilters.py:49-58):
 * FFT overlap-add with 512-tap minimum-phase FIR approximations of the same designs
 * (friture/filter.py:136-247, taps friture/data/generated_fft.npz).  Overlap-add is an exact linear
 * convolution, which the kernel computes directly (double accumulation), carrying the last
 * n_taps-1 inputs of every stage.  Parity mode; the IIR bank above is the throughput path.
 *   fir_band [bands_per_octave][n_taps], fir_dec [n_taps]  (float64 taps)
 *   process: n_samples <= 4096 and a multiple of 2^(n_octaves-1); y_dev ragged as in frt_bank_process */
int frt_firbank_plan(frt_handle h, int n_channels, int bands_per_octave, int n_octaves, int n_taps,
                     const double *fir_band, const double *fir_dec);
int frt_firbank_reset(frt_handle h);
int frt_firbank_process(frt_handle h, const float *x_dev, int64_t x_stride, int n_samples,
                        float *y_dev, int64_t y_stride, void *stream);

/* ---------------------------------------------------------------- combined per-hop analysis
 * What Spectrogram_Widget.handle_new_data (friture/spectrogram.py:131-169) and
 * OctaveSpectrum_Widget.handle_new_data (friture/octavespectrum.py:91-121) compute from the same
 * chunk of new samples, for n_channels independent streams held in HOST memory: the log-power
 * column of every hop (frt_stft_plan'ed size, frames as in frt_stft_process_host) and the smoothed
 * band levels after every hop-sized block (frt_bank_plan'ed bank, dB + weighting when db != 0).
 * H2D copy, the two kernels and the D2H copies are pipelined over time segments inside the call
 * (bench.py's `e2e`).  n_samples % hop == 0, hop a valid filterbank block;
 *   spec_host  [n_channels][(n_samples - n_fft)/hop + 1][n_fft/2 + 1]
 *   bands_host [n_channels][n_samples/hop][nbands]                                               */
int frt_combined_process_host(frt_handle h, const float *x_host, int64_t x_stride, int n_channels,
                              int64_t n_samples, int hop, float *spec_host, float *bands_host,
                              int nbands, int db);

/* ---------------------------------------------------------------- NVLink peer memory
 * Transport of the north-star's final all-gather of spectrogram columns: every process of the box
 * allocates its gathered buffer here (a whole cudaMalloc allocation, exportable over CUDA IPC),
 * exchanges the 64-byte handles, opens its peers' buffers and pushes its own columns into them with
 * the copy engines (cudaMemcpyAsync over NVLink/NVSwitch): no SM takes part, so the transfers do
 * not compete with the filterbank kernel for issue slots (friture_b200/peer.py).                   */
int frt_peer_alloc(frt_handle h, size_t bytes, void **out);
int frt_peer_free(frt_handle h, void *p);
int frt_peer_export(frt_handle h, void *p, void *handle64);            /* cudaIpcGetMemHandle  */
int frt_peer_import(frt_handle h, const void *handle64, void **out);   /* cudaIpcOpenMemHandle */
int frt_peer_close(frt_handle h, void *peer_ptr);
int frt_peer_copy(frt_handle h, void *dst, const void *src, size_t bytes, void *stream);
/* One-hop push of a block into the same offset of n_peers (<= 15) opened peer buffers by a small copy
 * kernel (n_ctas CTAs, 0 = 64): the block is read once from local HBM and stored to every peer
 * over NVLink with 128-bit stores.  Faster than the copy engines at 8 GPUs (not at 2 or 4).                  */
int frt_peer_push(frt_handle h, const void *src, void *const *peer_dst, int n_peers, size_t bytes,
                  int n_ctas, void *stream);

/* ---------------------------------------------------------------- GCC-PHAT (delay estimator)
 * Stands behind `generalized_cross_correlation(d0, d1)` (friture/signal/correlation.py:24-43)
 * and the smoothing + peak pick around it (friture/delay_estimator.py:129-142).               */

/* Plan for frames of `length` samples (even, 2^a 3^b 5^c, <= 28160; the widget's default is
 * 2 * delayrange_s * 12 kHz = 24000, delay_estimator.py:53-54,114-117).                       */
int frt_gcc_plan(frt_handle h, int length);
/* n_pairs independent channel pairs: d0/d1[p*stride + n].  Inputs are not modified (the
 * reference subtracts the means in place, a side effect that is not reproduced).
 *   xcorr_dev     NULL or [n_pairs][length]: this frame's Xcorr (correlation.py:41)
 *   smoothed_dev  NULL or [n_pairs][length]: smoothed Xcorr (0.3*X + 0.7*old, delay_estimator.py:134-139),
 *                 written for every non-silent pair.  have_prev: 0 = first frame, nothing to blend
 *                 (a silent pair is marked "no previous" with a NaN in its slot 0); 1 = blend every
 *                 pair with the buffer; 2 = blend the pairs that have had a non-silent frame (no
 *                 NaN marker) -- the widget's `old_Xcorr is not None`; silent pairs keep their buffer
 *   idx_dev/val_dev [n_pairs]: i = argmax |Xs| and Xs[i] (delay_estimator.py:141-146); pairs
 *                 with a constant input (std == 0) give (0, 0) as in delay_estimator.py:129-131,164-167 */
int frt_gcc_phat(frt_handle h, const float *d0_dev, const float *d1_dev, int64_t stride,
                 int n_pairs, float *xcorr_dev, float *smoothed_dev, int have_prev, int *idx_dev,
                 float *val_dev, void *stream);

/* ---------------------------------------------------------------- N-fold decimation
 * Stands behind `decimate_multiple(Ndec, bdec, adec, x, zis)` (friture/signal/decimate.py:45-71),
 * the 48 kHz -> 12 kHz pre-stage of the delay estimator (friture/delay_estimator.py:53-59,97-98):
 * n_stages x (order-12 elliptic low-pass as 6 float32 sections, keep even samples), state carried
 * across calls.  n_samples % 256 == 0; out holds n_samples >> n_stages samples per channel.      */
int frt_decimate_plan(frt_handle h, int n_channels, int n_stages, const double *sos_dec);
int frt_decimate_process(frt_handle h, const float *x_dev, int64_t x_stride, int64_t n_samples,
                         float *out_dev, int64_t out_stride, void *stream);

/* ---------------------------------------------------------------- spectrum widget reductions
 * The per-tick numeric work of Spectrum_Widget.handle_new_data behind the STFT
 * (friture/spectrum.py:158-181): smoothing across the tick's frames
 * (exp_smoothed_value_2d, friture/signal/exp_smoothing.py:59-107, alpha from spectrum.py:196-218),
 * 10*log10(sp+1e-30) + weighting (spectrum.py:95-101,171), arg-max (spectrum.py:175) and the
 * 3-harmonic product spectrum arg-max (spectrum.py:103-123,179-181).
 *   power_dev [C][n_frames][nbins] (strides in floats), disp_dev [C][nbins] smoothed power in/out
 *   (the widget's dispbuffers1), weight_dev [nbins] dB offsets or NULL, db_dev [C][nbins],
 *   fmax_idx_dev / pitch_idx_dev [C].                                                          */
int frt_spectrum_reduce(frt_handle h, const float *power_dev, int64_t stride_c, int64_t stride_f,
                        int n_channels, int n_frames, int nbins, float alpha, float *disp_dev,
                        const float *weight_dev, float *db_dev, int *fmax_idx_dev,
                        int *pitch_idx_dev, void *stream);

/* ---------------------------------------------------------------- spectrogram display chain
 * The Transform_Pipeline behind the spectrogram (friture/spectrogram.py:161-169): scaling of
 * dB + weighting to [0,1] (spectrogram.py:127-129), Frequency_Resampler.push
 * (friture/signal/frequency_resampler.py:67-83), Online_Linear_2D_resampler.push
 * (friture/signal/online_linear_2D_resampler.py:61-97) and Color_Transform.push
 * (friture/signal/color_tranform.py:48-51), fused.  The host supplies the screen-row table
 * (row_i0/row_t: bin index and fraction of every screen row on the chosen frequency scale) and,
 * per output column, the input column it is drawn from and the weight of the previous input
 * column (the resampler's index bookkeeping).  db_dev [C][n_frames][nbins] log-power columns,
 * old_data_dev [C][height] carried last column (in/out), lut_dev [256] 0xAARRGGBB,
 * pixels_dev [C][height][n_out].                                                              */
int frt_display_columns(frt_handle h, const float *db_dev, int64_t stride_c, int64_t stride_f,
                        int n_channels, int n_frames, int nbins, const float *weight_dev,
                        float spec_min, float spec_max, const int *row_i0_dev,
                        const float *row_t_dev, int height, const int *out_col_dev,
                        const float *out_a_dev, int n_out, float *old_data_dev,
                        const uint32_t *lut_dev, uint32_t *pixels_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FRT_H */

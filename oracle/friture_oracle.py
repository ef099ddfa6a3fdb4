"""NumPy/SciPy restatement of the reference's per-chunk spectral hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference file:line it restates (paths relative to the reference root).  All arithmetic
is float64, like the reference (friture/audiobackend.py:466-468 widens the device's
float32 to float64 before anything else touches it).

The FFTs are ``numpy.fft`` (pocketfft), i.e. the very third-party routine the reference
calls (friture/audioproc.py:23,44; friture/signal/correlation.py:21,34-41); the IIR
recursions are restated as a direct-form-II-transposed loop, run either through
``scipy.signal.lfilter`` (same recursion, verified bit-identical to the reference's
pure-Python loop in tests/test_oracle_vs_reference.py) or through the plain-C loop in
``oracle/iir_df2t.c``.
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 48000      # friture/audiobackend.py:31
FRAMES_PER_BUFFER = 512    # friture/audiobackend.py:32
NOCTAVE = 9                # friture/filter.py:7


# --------------------------------------------------------------------------- STFT
def hann_window(n_fft: int) -> np.ndarray:
    """Symmetric Hann, 0.5*(1-cos(2*pi*n/(N-1))): friture/audioproc.py:76-81."""
    n = np.arange(0, n_fft)
    return 0.5 * (1.0 - np.cos(2 * np.pi * n / (n_fft - 1)))


def analyzelive(samples: np.ndarray, window: np.ndarray | None = None) -> np.ndarray:
    """|rfft(x*w)|^2 / N^2: friture/audioproc.py:42-50,73-74."""
    samples = np.asarray(samples, dtype=np.float64)
    n_fft = samples.shape[-1]
    if window is None:
        window = hann_window(n_fft)
    fft = np.fft.rfft(samples * window)
    return (fft * fft.conjugate()).real / float(n_fft) ** 2


def log_spectrogram(sp: np.ndarray) -> np.ndarray:
    """10*log10(sp + 1e-30): friture/spectrogram.py:119-125, friture/spectrum.py:95-101."""
    return 10.0 * np.log10(sp + 1e-30)


def frame_count(n_samples: int, n_fft: int, hop: int) -> int:
    """Number of whole frames in a stream of n_samples (first frame = samples [0, n_fft))."""
    if n_samples < n_fft:
        return 0
    return (n_samples - n_fft) // hop + 1


def stft_power(x: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """Framing loop of friture/spectrogram.py:131-159 (spectrum.py:125-155) for one stream.

    Frame i is the n_fft samples ending at ``n_fft + i*hop`` (the reference's
    ``data_indexed(old_index, fft_size)`` returns the samples *ending* at old_index,
    friture/ringbuffer.py:87-99, and ``old_index += int(needed)`` per column).  Returns
    power[frames, n_fft//2+1] (time-major; the reference stores the transpose,
    spectrogram.py:147,157)."""
    x = np.asarray(x, dtype=np.float64)
    w = hann_window(n_fft)
    nf = frame_count(x.shape[-1], n_fft, hop)
    out = np.empty((nf, n_fft // 2 + 1), dtype=np.float64)
    for i in range(nf):
        out[i] = analyzelive(x[i * hop: i * hop + n_fft], w)
    return out


def stft_power_batch(x: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """Vectorised form of :func:`stft_power` for x[C, T] (same arithmetic, one rfft call)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    nf = frame_count(x.shape[-1], n_fft, hop)
    w = hann_window(n_fft)
    idx = np.arange(nf)[:, None] * hop + np.arange(n_fft)[None, :]
    fft = np.fft.rfft(x[:, idx] * w, axis=-1)
    return (fft * fft.conjugate()).real / float(n_fft) ** 2


def weighting_tables(f: np.ndarray, eps: float = 1e-50):
    """A/B/C psychoacoustic weighting in dB: friture/audioproc.py:83-96 (eps=1e-50);
    friture/octavefilters.py:76-82 uses the same formulas with no eps (pass eps=0)."""
    f = np.asarray(f, dtype=np.float64)
    Rc = 12200. ** 2 * f ** 2 / ((f ** 2 + 20.6 ** 2) * (f ** 2 + 12200. ** 2))
    Rb = 12200. ** 2 * f ** 3 / ((f ** 2 + 20.6 ** 2) * (f ** 2 + 12200. ** 2) * ((f ** 2 + 158.5 ** 2) ** 0.5))
    Ra = 12200. ** 2 * f ** 4 / ((f ** 2 + 20.6 ** 2) * (f ** 2 + 12200. ** 2) * ((f ** 2 + 107.7 ** 2) ** 0.5) * ((f ** 2 + 737.9 ** 2) ** 0.5))
    C = 0.06 + 20. * np.log10(Rc + eps)
    B = 0.17 + 20. * np.log10(Rb + eps)
    A = 2.0 + 20. * np.log10(Ra + eps)
    return A, B, C


# --------------------------------------------------------------------------- IIR
def lfilter_df2t(b, a, x, zi):
    """Direct-form-II-transposed IIR with carried state: friture/signal/lfilter.py:85-147.

    ``scipy.signal.lfilter`` runs the identical recursion (a[0]==1 in every reference
    coefficient set, friture/generated_filters.py), bit-for-bit equal to the reference's
    Python loop (checked in tests/test_oracle_vs_reference.py)."""
    from scipy.signal import lfilter
    y, zf = lfilter(np.asarray(b, dtype=np.float64), np.asarray(a, dtype=np.float64),
                    np.asarray(x, dtype=np.float64), zi=np.asarray(zi, dtype=np.float64))
    return y, zf


def lfilter_df2t_loop(b, a, x, zi):
    """Literal scalar loop of friture/signal/lfilter.py:131-139 (slow; small cases only)."""
    b = [float(v) for v in b]
    a = [float(v) for v in a]
    z = [float(v) for v in zi]
    nb = len(b)
    y = np.empty(len(x), dtype=np.float64)
    for k, xk in enumerate(np.asarray(x, dtype=np.float64).tolist()):
        yk = z[0] + b[0] * xk
        y[k] = yk
        for n in range(nb - 2):
            z[n] = z[n + 1] + xk * b[n + 1] - yk * a[n + 1]
        z[nb - 2] = xk * b[nb - 1] - yk * a[nb - 1]
    return y, np.array(z)


def decimate(bdec, adec, x, zi=None):
    """Low-pass then keep even indices: friture/signal/decimate.py:27-42."""
    if len(x) == 0:
        raise Exception("Filter input is too small")
    if zi is None:
        zi = np.zeros(max(len(bdec), len(adec)) - 1)
    y, zf = lfilter_df2t(bdec, adec, x, zi)
    return y[::2], zf


def decimate_multiple(ndec, bdec, adec, x, zis):
    """N-fold decimation with carried states: friture/signal/decimate.py:45-71."""
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0:
        return x, zis
    zfs = []
    for i in range(ndec):
        x, zf = decimate(bdec, adec, x, None if zis is None else zis[i])
        zfs.append(zf)
    return x, (None if zis is None else zfs)


def bank_filtic(bdec, adec, boct, aoct, noctave=NOCTAVE):
    """Zero initial states, ordered band(bpo-1..0) then decimator per stage:
    friture/filter.py:121-133."""
    zis = []
    for _ in range(noctave):
        for i in range(len(boct))[::-1]:
            zis.append(np.zeros(max(len(boct[i]), len(aoct[i])) - 1))
        zis.append(np.zeros(max(len(bdec), len(adec)) - 1))
    return zis


def octave_filter_bank_decimation(bdec, adec, boct, aoct, x, zis, noctave=NOCTAVE):
    """Multi-rate IIR bank, the numerical oracle of the filterbank row:
    friture/filter.py:86-118.  Returns (y[nbands] ragged, dec[nbands], zfs)."""
    bpo = len(boct)
    nb = noctave * bpo
    y = [None] * nb
    dec = [0] * nb
    zfs = []
    x_dec = np.asarray(x, dtype=np.float64)
    m = 0
    k = nb - 1
    for j in range(noctave):
        for i in range(bpo)[::-1]:
            filt, zf = lfilter_df2t(boct[i], aoct[i], x_dec, zis[m])
            m += 1
            zfs.append(zf)
            y[k] = filt
            dec[k] = 2 ** j
            k -= 1
        x_dec, zf = decimate(bdec, adec, x_dec, zis[m])
        m += 1
        zfs.append(zf)
    return y, dec, zfs


def get_decs(bpo, noctave=NOCTAVE):
    """friture/octavefilters.py:60-63."""
    return [2 ** j for j in range(noctave)[::-1] for _ in range(bpo)]


def octave_frequencies(total_bands_count, bands_per_octave):
    """Band centre / edge frequencies: friture/filter.py:39-54."""
    f0 = 1000.
    b = 1. / bands_per_octave
    imax = total_bands_count // 2
    if total_bands_count % 2 == 0:
        i = np.arange(-imax, imax)
    else:
        i = np.arange(-imax, imax + 1)
    fi = f0 * 2 ** (i * b)
    return fi, fi * np.sqrt(2 ** (-b)), fi * np.sqrt(2 ** b)


# --------------------------------------------------------------------------- live FFT-OLA bank
FIR_LENGTH = 512           # friture/octavefilters.py:35


def fft_bank_sizes(fir_length=FIR_LENGTH, noctave=NOCTAVE, max_block=1024):
    """Per-stage FFT sizes of the live bank: the next 2^a 3^b 5^c >= N_stage + fir_length - 1 with
    N_stage = max_block / 2^j (friture/filter_design.py:400-401 committed as
    friture/data/generated_fft.npz `*_fft_sizes` = [1536, 1024, 768, 640, 576, 576, 540, 540, 540])."""
    def next_composite(n):
        while True:
            m = n
            for p in (2, 3, 5):
                while m % p == 0:
                    m //= p
            if m == 1:
                return n
            n += 1
    return [next_composite((max_block >> j) + fir_length - 1) for j in range(noctave)]


def fft_bank_state(bpo, fir_length=FIR_LENGTH, noctave=NOCTAVE):
    """Zero overlap buffers (friture/octavefilters.py:145-158)."""
    return ([np.zeros((bpo, fir_length - 1)) for _ in range(noctave)],
            [np.zeros(fir_length - 1) for _ in range(noctave)])


def octave_filter_bank_decimation_fft(boct_fir, bdec_fir, x, overlaps_oct, overlaps_dec,
                                      noctave=NOCTAVE, fft_sizes=None):
    """The reference's live filterbank (friture/filter.py:136-247): per stage one shared rfft of
    the stage input, multiplication with the rfft of the 512-tap FIRs, irfft, pending overlap tail
    added element-wise (fixed fir_length-1 buffer), first N_s samples out, decimator output
    [:N_s:2] to the next stage.  Returns (y, dec, overlaps_oct, overlaps_dec) like the reference."""
    boct_fir = np.asarray(boct_fir, dtype=np.float64)
    bdec_fir = np.asarray(bdec_fir, dtype=np.float64)
    bpo, L = boct_fir.shape
    Lm1 = L - 1
    if fft_sizes is None:
        fft_sizes = fft_bank_sizes(L, noctave)
    nb = noctave * bpo
    y = [None] * nb
    dec = [0] * nb
    k = nb - 1
    x_dec = np.asarray(x, dtype=np.float64)
    ov_o, ov_d = [], []
    for j in range(noctave):
        N_s = len(x_dec)
        n = fft_sizes[j]
        X = np.fft.rfft(x_dec, n)
        y_oct = np.fft.irfft(X[None, :] * np.fft.rfft(boct_fir, n, axis=1), n, axis=1)
        y_dec = np.fft.irfft(X * np.fft.rfft(bdec_fir, n), n)
        pend, pend_d = overlaps_oct[j], overlaps_dec[j]
        add = min(pend.shape[1], N_s)
        y_oct[:, :add] += pend[:, :add]
        y_dec[:add] += pend_d[:add]
        for i in range(bpo)[::-1]:
            y[k] = y_oct[i, :N_s]
            dec[k] = 2 ** j
            k -= 1
        x_dec = y_dec[:N_s:2]
        tail = y_oct[:, N_s:N_s + Lm1].copy()
        rest = pend[:, add:]
        tail[:, :rest.shape[1]] += rest
        ov_o.append(tail)
        tail_d = y_dec[N_s:N_s + Lm1].copy()
        rest_d = pend_d[add:]
        tail_d[:len(rest_d)] += rest_d
        ov_d.append(tail_d)
    return y, dec, ov_o, ov_d


def fir_bank_direct(boct_fir, bdec_fir, x, hist, noctave=NOCTAVE):
    """The same filterbank as plain FIR convolutions with carried input history (what the GPU
    kernel computes): y_i = conv(x_j, h_i), x_{j+1} = conv(x_j, h_dec)[::2].  hist[j] holds the
    last L-1 inputs of stage j.  Equal to the overlap-add form up to rounding."""
    boct_fir = np.asarray(boct_fir, dtype=np.float64)
    bpo, L = boct_fir.shape
    nb = noctave * bpo
    y = [None] * nb
    k = nb - 1
    x_dec = np.asarray(x, dtype=np.float64)
    new_hist = []
    for j in range(noctave):
        N_s = len(x_dec)
        ext = np.concatenate([hist[j], x_dec])
        for i in range(bpo)[::-1]:
            y[k] = np.convolve(ext, boct_fir[i])[L - 1:L - 1 + N_s]
            k -= 1
        yd = np.convolve(ext, bdec_fir)[L - 1:L - 1 + N_s]
        new_hist.append(ext[-(L - 1):])
        x_dec = yd[::2]
    return y, new_hist


# --------------------------------------------------------------------------- smoothing
def smoothing_alpha(response_time, rate):
    """alpha so that the newest n = T*rate samples carry 65 % of the weight:
    friture/octavespectrum.py:140-156 (rate = fs/dec), friture/spectrum.py:196-218
    (rate = fs/hop)."""
    w = 0.65
    n = response_time * rate
    return 1. - (1. - w) ** (1. / (n + 1))


def smoothing_kernel(alpha, n):
    """(1-alpha)^(n-1 .. 0): friture/octavespectrum.py:75-79, friture/spectrum.py:220-222."""
    return (1. - alpha) ** np.arange(int(n) - 1, -1, -1)


def exp_smoothed_value(kernel, alpha, data, previous):
    """Block form of s <- alpha*x + (1-alpha)*s: friture/signal/exp_smoothing.py:11-56."""
    N = data.shape[0]
    Nk = kernel.shape[0]
    if N > Nk:
        N = Nk
        a = 0.0
    else:
        a = (1.0 - alpha) ** N
    if N == 0:
        return previous
    conv = np.dot(kernel[Nk - N:Nk], data[:N])
    return float(alpha * conv + previous * a)


def exp_smoothed_value_2d(kernel, alpha, data, previous):
    """Row-wise block smoothing: friture/signal/exp_smoothing.py:59-107."""
    Nf, Nt = data.shape
    Nk = kernel.shape[0]
    if Nt > Nk:
        Nt = Nk
        a = 0.0
    else:
        a = (1.0 - alpha) ** Nt
    if Nt == 0:
        return np.array(previous, copy=True)
    conv = data[:, :Nt] @ kernel[Nk - Nt:Nk]
    return alpha * conv + previous * a


class OctaveSpectrumOracle:
    """What OctaveSpectrum_Widget.handle_new_data does per chunk, with the IIR bank
    (friture/octavespectrum.py:91-121,140-156 around friture/filter.py:86-118):
    filter -> y**2 -> exp_smoothed_value per band -> 10*log10(sp+1e-30) (+ weighting)."""

    def __init__(self, bdec, adec, boct, aoct, response_time=1.0, noctave=NOCTAVE):
        self.bdec = np.asarray(bdec, dtype=np.float64)
        self.adec = np.asarray(adec, dtype=np.float64)
        self.boct = [np.asarray(v, dtype=np.float64) for v in boct]
        self.aoct = [np.asarray(v, dtype=np.float64) for v in aoct]
        self.noctave = noctave
        self.bpo = len(boct)
        self.nbands = noctave * self.bpo
        self.zis = bank_filtic(self.bdec, self.adec, self.boct, self.aoct, noctave)
        decs = get_decs(self.bpo, noctave)
        self.decs = decs
        self.alphas = [smoothing_alpha(response_time, SAMPLING_RATE / d) for d in decs]
        self.kernels = [smoothing_kernel(a, 2 * 4096 / d) for a, d in zip(self.alphas, decs)]
        self.dispbuffers = [0.0] * self.nbands

    def filter(self, x):
        y, dec, self.zis = octave_filter_bank_decimation(
            self.bdec, self.adec, self.boct, self.aoct, x, self.zis, self.noctave)
        return y, dec

    def push(self, x):
        """One chunk -> (smoothed band energies[nbands], dB[nbands], y)."""
        y, _ = self.filter(x)
        sp = [exp_smoothed_value(k, a, yy ** 2, old)
              for yy, k, a, old in zip(y, self.kernels, self.alphas, self.dispbuffers)]
        self.dispbuffers = sp
        sp = np.array(sp)
        return sp, 10 * np.log10(sp + 1e-30), y


# --------------------------------------------------------------------------- GCC-PHAT
def generalized_cross_correlation(d0, d1):
    """GCC-PHAT: friture/signal/correlation.py:24-43 (works on copies: the reference
    subtracts the means in place, a side effect that is not part of the result)."""
    d0 = np.array(d0, dtype=np.float64)
    d1 = np.array(d1, dtype=np.float64)
    d0 -= d0.mean()
    d1 -= d1.mean()
    window = np.hanning(len(d0))
    D0 = np.fft.rfft(d0 * window)
    D1 = np.fft.rfft(d1 * window)
    G = D0.conjugate() * D1
    absG = np.abs(G)
    m = max(absG)
    W = 1. / (1e-10 * m + absG)
    return np.fft.irfft(W * G)


def delay_peak(xcorr, old_xcorr=None, alpha=0.3):
    """Smoothing + peak pick of friture/delay_estimator.py:134-142.
    Returns (index, extremum value, smoothed xcorr)."""
    if old_xcorr is not None and old_xcorr.shape == xcorr.shape:
        sm = alpha * xcorr + (1. - alpha) * old_xcorr
    else:
        sm = xcorr
    i = int(np.argmax(np.abs(sm)))
    return i, float(sm[i]), sm


# --------------------------------------------------------------------------- spectrum widget reductions
def harmonic_product_spectrum(sp):
    """sp[:h]*sp[::2][:h]*sp[::3][:h], h=len//3: friture/spectrum.py:103-123."""
    h = sp.shape[0] // 3
    return sp[:h] * sp[::2][:h] * sp[::3][:h]


class SpectrumWidgetOracle:
    """Numeric part of Spectrum_Widget.handle_new_data for one channel
    (friture/spectrum.py:125-184 with setresponsetime :196-218): per tick, `analyzelive` on every
    realizable frame, exp_smoothed_value_2d across them, dB + weighting, arg-max, HPS arg-max."""

    def __init__(self, fft_size, overlap=0.75, response_time=0.025, weight=None):
        self.fft_size = fft_size
        self.hop = int(fft_size * (1. - overlap))
        n = response_time * SAMPLING_RATE / (fft_size * (1. - overlap))
        self.alpha = 1. - (1. - 0.65) ** (1. / (n + 1))
        self.kernel = smoothing_kernel(self.alpha, 2 * 4096)
        self.freq = np.linspace(0, SAMPLING_RATE // 2, fft_size // 2 + 1)
        self.w = np.zeros(fft_size // 2 + 1) if weight is None else weight
        self.disp = np.zeros(fft_size // 2 + 1)

    def tick(self, x):
        spn = stft_power(x, self.fft_size, self.hop).T           # [bins, realizable]
        sp = exp_smoothed_value_2d(self.kernel, self.alpha, spn, self.disp)
        self.disp = sp
        db = log_spectrogram(sp) + self.w
        i = int(np.argmax(db))
        pitch = int(np.argmax(harmonic_product_spectrum(sp)))
        return db, self.freq[i], max(self.freq[pitch], 1e-20), i, pitch


# --------------------------------------------------------------------------- display chain
def frequency_resample(data, freq, xscaled):
    """Frequency_Resampler.push: np.interp per column (friture/signal/frequency_resampler.py:67-83).
    data [bins, n] -> [len(xscaled), n]."""
    out = np.zeros((len(xscaled), data.shape[1]))
    for j in range(data.shape[1]):
        out[:, j] = np.interp(xscaled, freq, data[:, j])
    return out


class OnlineLinear2DResamplerOracle:
    """Online_Linear_2D_resampler (friture/signal/online_linear_2D_resampler.py:13-97 with
    linear_interp_2D, friture/signal/linear_interp.py:11-62), height fixed."""

    def __init__(self, interp_factor_L, decim_factor_M, height):
        self.ratio = float(interp_factor_L) / decim_factor_M
        self.orig_index = 0.
        self.resampled_index = 0.
        self.old_data = np.zeros(height)

    def processable(self, m):
        return int(np.ceil((self.orig_index + m - (self.resampled_index + self.ratio)) / self.ratio))

    def push(self, data):
        cols = []
        for j in range(data.shape[1]):
            self.orig_index += 1.
            n = self.processable(0)
            if n > 0:
                new_indices = self.resampled_index + self.ratio * np.arange(1, n + 1, dtype=np.float64)
                a = self.orig_index - new_indices
                cols.append(data[:, j][:, None] * (1.0 - a)[None, :] + self.old_data[:, None] * a[None, :])
                self.resampled_index = float(new_indices[-1])
            self.old_data = data[:, j]
        if not cols:
            return np.zeros((data.shape[0], 0))
        return np.concatenate(cols, axis=1)


def color_transform(lut, data):
    """Color_Transform.push: clip to [0,1], lut[int(v*255)]
    (friture/signal/color_tranform.py:48-51, friture/signal/lookup_table.py:32-52)."""
    return lut[(np.clip(data, 0., 1.) * 255).astype(np.intp)]

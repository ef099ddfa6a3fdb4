"""Numeric part of ``Spectrum_Widget`` (friture/spectrum.py:125-222) for many channels on the GPU:
framing + ``analyzelive`` per frame, exponential smoothing across the tick's frames, weighting,
dB, peak frequency and harmonic-product-spectrum pitch -- one smoothed column per channel per
tick instead of one column per frame (this is what the display consumes, and it shrinks what a
multi-GPU gather has to move from ``frames x bins`` to ``bins`` per channel)."""
from __future__ import annotations

from ctypes import c_float, c_void_p

import numpy as np

from . import _lib
from .audioproc import SAMPLING_RATE, audioproc, frame_count

DEFAULT_FFT_SIZE = 8192        # friture/spectrum_settings.py:27 (2**7 * 32 ... the widget default)
DEFAULT_RESPONSE_TIME = 0.025  # friture/spectrum_settings.py
OVERLAP = 0.75                 # friture/spectrum.py:66


class SpectrumAnalyzer:
    def __init__(self, n_channels, fft_size=DEFAULT_FFT_SIZE, overlap=OVERLAP,
                 response_time=DEFAULT_RESPONSE_TIME, weighting=0, handle=None):
        self.n_channels = int(n_channels)
        self.proc = audioproc(handle)
        self.overlap = overlap
        self.weighting = weighting
        self.response_time = response_time
        self._disp = None
        self._w_dev = None
        self.setfftsize(fft_size)

    # -- the widget's setters ---------------------------------------------------------------
    def setfftsize(self, fft_size):
        self.fft_size = fft_size
        self.proc.set_fftsize(fft_size)
        self.freq = self.proc.get_freq_scale()
        self.hop = int(self.fft_size * (1. - self.overlap))          # spectrum.py:137,155
        self.setresponsetime(self.response_time)
        self.setweighting(self.weighting)
        self._disp = None                                            # update_display_buffers, :224-226

    def setresponsetime(self, response_time):
        # spectrum.py:196-218: the newest n = T*fs/hop frames carry 65 % of the weight
        self.response_time = response_time
        w = 0.65
        n = self.response_time * SAMPLING_RATE / (self.fft_size * (1. - self.overlap))
        self.alpha = 1. - (1. - w) ** (1. / (n + 1))

    def setweighting(self, weighting):
        # 0 none, 1 A, 2 B, 3 C (spectrum.py:240-252)
        self.weighting = weighting
        A, B, C = self.proc.get_freq_weighting()
        self.w = [0. * A, A, B, C][weighting]
        self._w_dev = None

    # -- one tick ----------------------------------------------------------------------------
    def process(self, x, stream=None):
        """x: CUDA float32 [C, T] holding the tick's new frames (frame f = x[:, f*hop : f*hop+N]).
        Returns (dB [C, bins] CUDA tensor, fmax [C] Hz, fpitch [C] Hz) after smoothing across
        the frames, with the smoothing state carried from tick to tick."""
        import torch
        if x.dim() == 1:
            x = x[None, :]
        C = x.shape[0]
        if C != self.n_channels:
            raise ValueError("expected %d channels" % self.n_channels)
        nb = self.fft_size // 2 + 1
        nf = frame_count(x.shape[1], self.fft_size, self.hop)
        power = self.proc.stft(x, hop=self.hop, log=False, stream=stream)
        if self._disp is None:
            self._disp = torch.zeros((C, nb), dtype=torch.float32, device=x.device)
        if self._w_dev is None:
            self._w_dev = torch.from_numpy(np.ascontiguousarray(self.w, dtype=np.float32)).to(x.device)
        db = torch.empty((C, nb), dtype=torch.float32, device=x.device)
        imax = torch.empty(C, dtype=torch.int32, device=x.device)
        ipitch = torch.empty(C, dtype=torch.int32, device=x.device)
        sp = _lib.current_stream_ptr(x.device) if stream is None else c_void_p(int(stream))
        self.proc.handle.call("frt_spectrum_reduce", _lib._ptr(power), int(nf * nb), int(nb), int(C),
                              int(nf), int(nb), c_float(self.alpha), _lib._ptr(self._disp),
                              _lib._ptr(self._w_dev) if self.weighting else None, _lib._ptr(db),
                              _lib._ptr(imax), _lib._ptr(ipitch), sp)
        fmax = self.freq[imax.cpu().numpy()]
        fpitch = np.maximum(self.freq[ipitch.cpu().numpy()], 1e-20)     # spectrum.py:181
        return db, fmax, fpitch

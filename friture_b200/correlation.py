"""Drop-in for ``friture.signal.correlation.generalized_cross_correlation``
(friture/signal/correlation.py:24-43) and a batched delay estimator with the smoothing and peak
pick of ``Delay_Estimator_Widget.handle_new_data`` (friture/delay_estimator.py:129-152)."""
from __future__ import annotations

from ctypes import c_void_p

import numpy as np

from . import _lib
from ._lib import Handle, default_handle


def generalized_cross_correlation(d0, d1, handle=None):
    """GCC-PHAT of one pair: float64 NumPy in, float64 ``Xcorr[len(d0)]`` out.  Unlike the
    reference the inputs are left untouched (it subtracts the means in place)."""
    import torch
    d0 = np.ascontiguousarray(d0, dtype=np.float32)
    d1 = np.ascontiguousarray(d1, dtype=np.float32)
    if d0.ndim != 1 or d0.shape != d1.shape:
        raise ValueError("d0 and d1 must be 1-D arrays of the same length")
    est = GccPhat(d0.shape[0], handle=handle)
    dev = torch.device("cuda", est.handle.device)
    _, _, x = est.estimate(torch.from_numpy(d0).to(dev)[None, :], torch.from_numpy(d1).to(dev)[None, :],
                           smooth=False, want_xcorr=True)
    return x[0].cpu().numpy().astype(np.float64)


class GccPhat:
    """Batched GCC-PHAT over independent channel pairs.  ``estimate`` returns
    (index of max |Xs|, Xs at that index, Xcorr or None); with ``smooth`` the handle keeps
    ``Xs = 0.3*X + 0.7*Xs_prev`` per pair across calls like the widget's ``old_Xcorr``."""

    def __init__(self, length, handle=None, fs=12000.0):
        self.length = int(length)
        self.fs = fs                      # 48 kHz / 2**Ndec, delay_estimator.py:53-54
        self._handle = handle
        self._smoothed = None
        self._have_prev = False

    @property
    def handle(self):
        if self._handle is None:
            self._handle = default_handle()
        return self._handle

    def reset(self):
        self._smoothed = None
        self._have_prev = False

    def estimate(self, d0, d1, smooth=True, want_xcorr=False, stream=None):
        import torch
        if d0.dim() == 1:
            d0, d1 = d0[None, :], d1[None, :]
        if (d0.shape != d1.shape or d0.dtype != torch.float32 or d1.dtype != torch.float32
                or not d0.is_cuda or not d1.is_cuda):
            raise ValueError("d0, d1 must be CUDA float32 tensors of the same shape [P, L]")
        P, L = d0.shape
        if L != self.length:
            raise ValueError("frame length %d != planned %d" % (L, self.length))
        self.handle.check_device(d0)
        self.handle.check_device(d1)
        d0 = d0.contiguous()
        d1 = d1.contiguous()
        self.handle.call("frt_gcc_plan", int(L))
        idx = torch.empty(P, dtype=torch.int32, device=d0.device)
        val = torch.empty(P, dtype=torch.float32, device=d0.device)
        xc = torch.empty((P, L), dtype=torch.float32, device=d0.device) if want_xcorr else None
        sm = None
        have_prev = 0
        if smooth:
            if self._smoothed is None or tuple(self._smoothed.shape) != (P, L):
                self._smoothed = torch.zeros((P, L), dtype=torch.float32, device=d0.device)
                self._have_prev = False
            sm = self._smoothed
            # 2 = per pair: a pair whose frames have all been silent so far has no previous smoothed
            # frame (the kernel marks it), like old_Xcorr = None in the widget
            have_prev = 2 if self._have_prev else 0
        sp = _lib.current_stream_ptr(d0.device) if stream is None else c_void_p(int(stream))
        self.handle.call("frt_gcc_phat", _lib._ptr(d0), _lib._ptr(d1), int(L), int(P),
                         _lib._ptr(xc), _lib._ptr(sm), have_prev, _lib._ptr(idx), _lib._ptr(val), sp)
        if smooth:
            self._have_prev = True
        return idx, val, xc

    def delay_ms(self, idx):
        """Index -> delay in ms, wrapped to +-L/2 (delay_estimator.py:147-152)."""
        d = 1e3 * idx.to("cpu").double().numpy() / self.fs
        span = 1e3 * self.length / self.fs
        return np.where(d > span / 2.0, d - span, d)

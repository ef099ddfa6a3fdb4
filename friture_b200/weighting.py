"""A / B / C psychoacoustic weighting curves in dB (closed forms used by the reference in
friture/audioproc.py:88-96 for FFT bins, with eps = 1e-50, and friture/octavefilters.py:76-82 for
band centres, without eps)."""
from __future__ import annotations

import numpy as np

_F1, _F2, _F3, _F4, _F5 = 20.6, 107.7, 737.9, 12200., 158.5


def abc_weighting(f, eps: float = 0.0):
    """Return (A, B, C) in dB at frequencies `f` (Hz)."""
    f = np.asarray(f, dtype=np.float64)
    f2 = f ** 2
    low_high = (f2 + _F1 ** 2) * (f2 + _F4 ** 2)
    Rc = _F4 ** 2 * f ** 2 / low_high
    Rb = _F4 ** 2 * f ** 3 / (low_high * ((f2 + _F5 ** 2) ** 0.5))
    Ra = _F4 ** 2 * f ** 4 / (low_high * ((f2 + _F2 ** 2) ** 0.5) * ((f2 + _F3 ** 2) ** 0.5))
    C = 0.06 + 20. * np.log10(Rc + eps)
    B = 0.17 + 20. * np.log10(Rb + eps)
    A = 2.0 + 20. * np.log10(Ra + eps)
    return A, B, C

"""Filter coefficients of the fractional-octave bank.

``data/filters.npz`` carries the reference's committed designs (friture/generated_filters.py:5-884:
order-12 elliptic decimator, order-4 elliptic band-passes of the top octave for 1/3/6/12/24
bands per octave) as (b, a) and as second-order sections; ``tools/make_filter_data.py``
regenerates it from the reference tree.
"""
from __future__ import annotations

import os

import numpy as np

NOCTAVE = 9   # friture/filter.py:7
_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "filters.npz")
_cache = None

SUPPORTED_BPO = (1, 3, 6, 12, 24)   # friture/octavespectrum_settings.py:28


def load():
    global _cache
    if _cache is None:
        with np.load(_PATH) as d:
            _cache = {k: d[k] for k in d.files}
    return _cache


def decimator():
    """(bdec, adec, sos_dec): PARAMS['dec'] of the reference."""
    d = load()
    return d["bdec"], d["adec"], d["sos_dec"]


def bands(bpo: int):
    """(boct[bpo,5], aoct[bpo,5], sos[bpo,2,6]): PARAMS[str(bpo)] of the reference."""
    if bpo not in SUPPORTED_BPO:
        raise Exception("Unknown bandsperoctave: %d" % (bpo))   # friture/octavefilters.py:103
    d = load()
    return d["b%d" % bpo], d["a%d" % bpo], d["sos%d" % bpo]


# ISO 3 preferred numbers, R80 series (friture/renard.py builds the same series bottom-up)
_R80 = [1.00, 1.03, 1.06, 1.09, 1.12, 1.15, 1.18, 1.22, 1.25, 1.28, 1.32, 1.36, 1.40, 1.45, 1.50,
        1.55, 1.60, 1.65, 1.70, 1.75, 1.80, 1.85, 1.90, 1.95, 2.00, 2.06, 2.12, 2.18, 2.24, 2.30,
        2.36, 2.43, 2.50, 2.58, 2.65, 2.72, 2.80, 2.90, 3.00, 3.07, 3.15, 3.25, 3.35, 3.45, 3.55,
        3.65, 3.75, 3.87, 4.00, 4.12, 4.25, 4.37, 4.50, 4.62, 4.75, 4.87, 5.00, 5.15, 5.30, 5.45,
        5.60, 5.80, 6.00, 6.15, 6.30, 6.50, 6.70, 6.90, 7.10, 7.30, 7.50, 7.75, 8.00, 8.25, 8.50,
        8.75, 9.00, 9.25, 9.50, 9.75]


def renard(n: int):
    """R10 / R20 / R40 / R80 preferred numbers (every (80/n)-th term of R80)."""
    if n not in (10, 20, 40, 80):
        raise ValueError("no Renard series R%d" % n)
    return _R80[::80 // n]


_FIR_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "fir.npz")
_fir_cache = None


def fir_taps(bpo: int):
    """(boct_fir[bpo, 512], bdec_fir[512]): the 512-tap minimum-phase FIR approximations the
    reference's live filterbank runs by FFT overlap-add (friture/data/generated_fft.npz, loaded at
    friture/octavefilters.py:123-143); ``tools/make_fir_data.py`` regenerates ``data/fir.npz``."""
    global _fir_cache
    if bpo not in SUPPORTED_BPO:
        raise Exception("Unknown bandsperoctave: %d" % (bpo))
    if _fir_cache is None:
        with np.load(_FIR_PATH) as d:
            _fir_cache = {k: d[k] for k in d.files}
    return _fir_cache["boct_fir%d" % bpo], _fir_cache["bdec_fir"]

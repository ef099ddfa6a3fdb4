"""ctypes wrapper of oracle/_build/liboracle_iir.so (plain-C float64 IIR restatement).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "liboracle_iir.so")
_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        if not os.path.isfile(_PATH) and build:
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = ctypes.CDLL(_PATH)
        _lib.frt_oracle_bank_stream.restype = ctypes.c_int
        _lib.frt_oracle_bank_block.restype = ctypes.c_int
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def lfilter(b, a, x, zi):
    lib = load()
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    z = np.array(zi, dtype=np.float64)
    y = np.empty_like(x)
    lib.frt_oracle_lfilter(_p(b), _p(a), ctypes.c_int(len(b)), _p(x), ctypes.c_long(len(x)),
                           _p(z), _p(y))
    return y, z


class BankC:
    """Many channels x many blocks through the C bank; carries state like the widget."""

    def __init__(self, bdec, adec, boct, aoct, alphas_per_band, n_channels, noctave=9):
        self.lib = load()
        self.bdec = np.ascontiguousarray(bdec, dtype=np.float64)
        self.adec = np.ascontiguousarray(adec, dtype=np.float64)
        self.boct = np.ascontiguousarray(np.stack(boct), dtype=np.float64)
        self.aoct = np.ascontiguousarray(np.stack(aoct), dtype=np.float64)
        self.bpo = self.boct.shape[0]
        self.noct = noctave
        self.nbands = self.bpo * noctave
        self.alphas = np.ascontiguousarray(alphas_per_band, dtype=np.float64)
        self.C = n_channels
        self.zstride = noctave * (self.bpo * 4 + 12)
        self.zis = np.zeros((n_channels, self.zstride))
        self.energies = np.zeros((n_channels, self.nbands))

    def process(self, x, block, c0=0, c1=None):
        """x float32 [C, n_blocks*block] -> energies after each block [c1-c0, n_blocks, nbands]."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        c1 = self.C if c1 is None else c1
        nb = x.shape[1] // block
        out = np.zeros((self.C, nb, self.nbands))
        rc = self.lib.frt_oracle_bank_stream(
            _p(self.bdec), _p(self.adec), ctypes.c_int(len(self.bdec)), _p(self.boct),
            _p(self.aoct), ctypes.c_int(self.boct.shape[1]), ctypes.c_int(self.bpo),
            ctypes.c_int(self.noct), _p(x), ctypes.c_long(x.shape[1]), ctypes.c_int(c0),
            ctypes.c_int(c1), ctypes.c_long(block), ctypes.c_int(nb), _p(self.zis),
            ctypes.c_long(self.zstride), _p(self.alphas), _p(self.energies), _p(out))
        if rc:
            raise RuntimeError("frt_oracle_bank_stream failed: %d" % rc)
        return out[c0:c1]

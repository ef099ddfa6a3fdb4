#!/usr/bin/env python
"""Generate friture_b200/data/fir.npz from the reference's committed FIR designs.

The live filterbank of the reference (friture/octavefilters.py:123-158, friture/filter.py:136-247)
runs 512-tap minimum-phase FIR approximations of the IIR designs by FFT overlap-add; the taps are
design output committed as friture/data/generated_fft.npz (`bdec_fir`, `{bpo}_boct_fir`; the
`*_fft_H_*` arrays are their rffts at the per-stage FFT sizes).  Only the time-domain taps are
carried over: an exact FIR convolution with them IS what the overlap-add computes.  Needs
/root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    src = os.path.join(ref_import.REFERENCE_ROOT, "friture", "data", "generated_fft.npz")
    out = {}
    with np.load(src) as d:
        out["bdec_fir"] = np.asarray(d["bdec_fir"], dtype=np.float64)
        for bpo in (1, 3, 6, 12, 24):
            out["boct_fir%d" % bpo] = np.asarray(d["%d_boct_fir" % bpo], dtype=np.float64)
            # the frequency responses must be the rfft of these very taps (checked, not stored)
            sizes = d["%d_fft_sizes" % bpo]
            H = d["%d_fft_H_oct" % bpo]
            for j, n in enumerate(sizes):
                ref = np.fft.rfft(out["boct_fir%d" % bpo], int(n), axis=1)
                assert np.allclose(H[j][:, :int(n) // 2 + 1], ref, rtol=0, atol=1e-12), (bpo, j)
            Hd = d["%d_fft_H_dec" % bpo]
            for j, n in enumerate(sizes):
                assert np.allclose(Hd[j][:int(n) // 2 + 1], np.fft.rfft(out["bdec_fir"], int(n)), rtol=0, atol=1e-12)
    path = os.path.join(ROOT, "friture_b200", "data", "fir.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

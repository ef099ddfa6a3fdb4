#!/usr/bin/env python
"""Generate friture_b200/data/filters.npz from the reference's committed IIR designs.

The reference's coefficient source of truth is the (b, a) tables of
friture/generated_filters.py:5-884 (designed offline by friture/filter_design.py:131-164,181 with
an older SciPy -- re-running the design with today's scipy.signal.ellip gives coefficients that
differ by ~4e-5, so the numbers themselves must be carried over for parity).  This script reads
them through the reference's own module (oracle/ref_import.py; needs /root/reference) and writes
  bdec, adec            (13,)      order-12 elliptic decimation low-pass
  sos_dec               (6, 6)     the same filter as second-order sections (float64)
  b{bpo}, a{bpo}        (bpo, 5)   order-4 elliptic band-passes of the top octave
  sos{bpo}              (bpo, 2, 6)
for bpo in 1, 3, 6, 12, 24.  SOS = scipy.signal.tf2sos (pairing 'nearest'); the float64 SOS
cascade reproduces the (b, a) recursion to 2e-11, the float32 cascade to ~2e-6 (checked in
tests/test_filter_data.py).
"""
import os
import sys

import numpy as np
from scipy.signal import tf2sos

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    ref = ref_import.load()
    P = ref.generated_filters.PARAMS
    out = {}
    bdec, adec = (np.asarray(v, dtype=np.float64) for v in P["dec"])
    out["bdec"], out["adec"] = bdec, adec
    out["sos_dec"] = tf2sos(bdec, adec)
    for bpo in (1, 3, 6, 12, 24):
        boct, aoct = P[str(bpo)][0], P[str(bpo)][1]
        b = np.asarray(boct, dtype=np.float64)
        a = np.asarray(aoct, dtype=np.float64)
        out["b%d" % bpo], out["a%d" % bpo] = b, a
        out["sos%d" % bpo] = np.stack([tf2sos(b[i], a[i]) for i in range(bpo)])
    path = os.path.join(ROOT, "friture_b200", "data", "filters.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate friture_b200/data/cmrmap.npz: the 256-entry colour map of the spectrogram
(friture/plotting/generated_cmrmap.py, a generated table; Rappaport 2002 CMR map) packed as the
uint32 0xAARRGGBB words the reference builds with QColor(int(r*255), int(g*255), int(b*255)).rgb()
(friture/signal/color_tranform.py:35-45).  Needs /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    ref_import.load()
    from friture.plotting import generated_cmrmap
    cmap = np.asarray(generated_cmrmap.CMAP, dtype=np.float64)
    r = (cmap[:, 0] * 255).astype(np.int64)
    g = (cmap[:, 1] * 255).astype(np.int64)
    b = (cmap[:, 2] * 255).astype(np.int64)
    lut = (0xFF000000 | (r << 16) | (g << 8) | b).astype(np.uint32)
    path = os.path.join(ROOT, "friture_b200", "data", "cmrmap.npz")
    np.savez_compressed(path, lut=lut)
    print("wrote", path, lut.shape, hex(int(lut[0])), hex(int(lut[-1])))


if __name__ == "__main__":
    main()

"""CPU: the spectrogram display chain -- oracle restatement and host-side logic against golden
vectors produced by the reference's Frequency_Resampler / Online_Linear_2D_resampler classes."""
import os
from fractions import Fraction

import numpy as np

from oracle import friture_oracle as fo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    with np.load(os.path.join(GOLD, "display.npz")) as d:
        return {k: d[k] for k in d.files}


def ratio_LM():
    return Fraction(48000, 2048) / (Fraction(1) - Fraction(3, 4)) / 1000, Fraction(700, 10000)


def test_oracle_display_chain_matches_reference():
    g = load()
    freq = np.linspace(0, 24000, 1025)
    L, M = ratio_LM()
    tr = fo.OnlineLinear2DResamplerOracle(L, M, 96)
    for tick in range(3):
        norm = (g["db_%d" % tick].astype(np.float64) + 140.0) / 140.0
        res = tr.push(fo.frequency_resample(norm, freq, g["xscaled"]))
        assert res.shape == g["resampled_%d" % tick].shape
        assert np.array_equal(res, g["resampled_%d" % tick])
        assert np.array_equal(fo.color_transform(g["lut"], res), g["pixels_%d" % tick])


def test_host_tables_and_index_bookkeeping():
    from friture_b200.display import Mel, OnlineResamplerIndex, load_lut, screen_rows, SCALES
    g = load()
    freq = np.linspace(0, 24000, 1025)
    xs, i0, t = screen_rows(freq, Mel, 20., 24000., 96)
    assert np.array_equal(xs, g["xscaled"])
    col = np.random.default_rng(0).random(1025)
    assert np.allclose(col[i0] + t * (col[i0 + 1] - col[i0]), np.interp(xs, freq, col), rtol=1e-13)
    assert np.array_equal(load_lut(), g["lut"])
    L, M = ratio_LM()
    idx = OnlineResamplerIndex()
    idx.set_ratio(L, M)
    assert abs(idx.resampling_ratio - float(g["ratio"])) < 1e-15
    for tick, ncols in enumerate((5, 1, 9)):
        cols, a = idx.push(ncols)
        assert len(cols) == g["resampled_%d" % tick].shape[1]
        assert np.all((a >= -1e-12) & (a <= 1 + 1e-12)) and np.all(np.diff(cols) >= 0)
    for s in SCALES.values():       # transform / inverse are inverse pairs
        f = np.array([20., 440., 1000., 20000.])
        assert np.allclose(s.inverse(s.transform(f)), f, rtol=1e-12)

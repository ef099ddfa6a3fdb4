"""CPU: host-side logic of the drop-in classes against the golden vectors (no GPU needed)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLD, name)) as d:
        return {k: d[k] for k in d.files}


def test_audioproc_attributes():
    from friture_b200 import audioproc
    from friture_b200.audioproc import frame_count
    g = load("analyzelive.npz")
    p = audioproc()
    assert p.fft_size == 10 and len(p.freq) == 10      # initial state of the reference class
    for n in (1024, 2048, 8192):
        p.set_fftsize(n)
        assert np.array_equal(p.window, g["window_%d" % n])
        assert np.array_equal(p.get_freq_scale(), g["freq_%d" % n])
        A, B, C = p.get_freq_weighting()
        assert np.array_equal(A, g["A_%d" % n]) and np.array_equal(B, g["B_%d" % n])
        assert np.array_equal(C, g["C_%d" % n])
        assert p.size_sq == float(n) ** 2
    assert frame_count(2047, 2048, 1024) == 0
    assert frame_count(2048, 2048, 1024) == 1
    assert frame_count(2048 + 1023, 2048, 1024) == 1
    assert frame_count(2048 + 1024, 2048, 1024) == 2


def test_octave_filters_host_side():
    from friture_b200.octavefilters import (Octave_Filters, ragged_layout, smoothing_alphas)
    from oracle import friture_oracle as fo
    g = load("octave_bank.npz")
    c = load("coefficients.npz")
    of = Octave_Filters(3)
    assert of.nbands == 27 and of.bandsperoctave == 3
    for name in ("fi", "flow", "fhigh", "A", "B", "C"):
        assert np.array_equal(getattr(of, name), g[name + "_bpo3"]), name
    assert of.f_nominal == list(g["f_nominal_bpo3"])
    assert of.get_decs() == list(g["dec_bpo3"])
    assert np.array_equal(of.bdec, c["bdec"]) and np.array_equal(np.stack(of.boct), c["b3"])
    for bpo in (1, 6, 12, 24):
        of.setbandsperoctave(bpo)
        assert of.nbands == 9 * bpo and len(of.f_nominal) == of.nbands
        assert np.array_equal(np.stack(of.aoct), c["a%d" % bpo])
    al = smoothing_alphas(1.0)
    assert np.allclose(al, [fo.smoothing_alpha(1.0, 48000 / 2 ** j) for j in range(9)], rtol=1e-15)
    off, ln = ragged_layout(512, 3)
    assert ln[:3] == [2, 2, 2] and ln[-3:] == [512] * 3 and off[0] == 0
    assert off[-1] + ln[-1] == sum(3 * (512 >> j) for j in range(9))


def test_sos_sections_reproduce_ba_filters():
    """float64 SOS == (b, a) recursion to ~1e-11; float32 SOS to ~1e-5."""
    from scipy.signal import sosfilt
    from friture_b200 import filter_data
    from oracle import friture_oracle as fo
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4096) * 0.1
    bdec, adec, sos = filter_data.decimator()
    ref, _ = fo.lfilter_df2t(bdec, adec, x, np.zeros(12))
    assert np.max(np.abs(sosfilt(sos, x) - ref)) / np.max(np.abs(ref)) < 1e-9
    got32 = sosfilt(sos.astype(np.float32), x.astype(np.float32))
    assert got32.dtype == np.float32
    assert np.max(np.abs(got32 - ref)) / np.max(np.abs(ref)) < 1e-5
    for bpo in filter_data.SUPPORTED_BPO:
        b, a, s = filter_data.bands(bpo)
        for i in range(bpo):
            ref, _ = fo.lfilter_df2t(b[i], a[i], x, np.zeros(4))
            assert np.max(np.abs(sosfilt(s[i], x) - ref)) / np.max(np.abs(ref)) < 1e-9

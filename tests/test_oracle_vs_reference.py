"""CPU, build container only: the oracle restatement against the UNMODIFIED reference imported in
place from /root/reference (skipped where the reference tree does not exist, e.g. the GPU box)."""
import numpy as np
import pytest

from oracle import friture_oracle as fo
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load()


def test_analyzelive_all_sizes(ref):
    for k in range(10):
        n = 32 * 2 ** k                                   # spectrum_settings.py:61-70
        rng = np.random.default_rng(n)
        x = rng.standard_normal(n)
        proc = ref.audioproc.audioproc()
        proc.set_fftsize(n)
        assert np.array_equal(proc.analyzelive(x.copy()), fo.analyzelive(x))
        assert np.array_equal(proc.window, fo.hann_window(n))


def test_lfilter_bit_identical(ref):
    from oracle import iir_c
    P = ref.generated_filters.PARAMS
    rng = np.random.default_rng(0)
    x = rng.standard_normal(700)
    for b, a in [(P["dec"][0], P["dec"][1]), (P["3"][0][1], P["3"][1][1])]:
        b, a = np.array(b), np.array(a)
        zi = rng.standard_normal(len(b) - 1) * 0.01
        y0, z0 = ref.lfilter.lfilter_float64_1D(b, a, x, zi.copy())
        for fn in (fo.lfilter_df2t, fo.lfilter_df2t_loop, iir_c.lfilter):
            y, z = fn(b, a, x, zi.copy())
            assert np.array_equal(y, y0) and np.array_equal(z, z0), fn


def test_bank_and_smoothing(ref):
    from friture.filter import octave_filter_bank_decimation, octave_filter_bank_decimation_filtic
    P = ref.generated_filters.PARAMS
    bdec, adec = np.array(P["dec"][0]), np.array(P["dec"][1])
    for bpo in (1, 3, 24):
        boct = [np.array(v) for v in P[str(bpo)][0]]
        aoct = [np.array(v) for v in P[str(bpo)][1]]
        rng = np.random.default_rng(bpo)
        zr = octave_filter_bank_decimation_filtic(bdec, adec, boct, aoct)
        zo = fo.bank_filtic(bdec, adec, boct, aoct)
        for _ in range(3):
            x = rng.standard_normal(512)
            yr, dr, zr = octave_filter_bank_decimation(bdec, adec, boct, aoct, x, zr)
            yo, do, zo = fo.octave_filter_bank_decimation(bdec, adec, boct, aoct, x, zo)
            assert dr == do
            assert all(np.array_equal(a, b) for a, b in zip(yr, yo))
            assert all(np.array_equal(a, b) for a, b in zip(zr, zo))
    k = fo.smoothing_kernel(0.01, 64)
    d = np.random.default_rng(1).random(40)
    assert ref.exp_smoothing.exp_smoothed_value(k, 0.01, d, 0.3) == fo.exp_smoothed_value(k, 0.01, d, 0.3)


def test_octave_filters_attributes_and_gcc(ref):
    of = ref.octavefilters.Octave_Filters(3)
    fi, fl, fh = fo.octave_frequencies(27, 3)
    assert np.array_equal(of.fi, fi) and np.array_equal(of.flow, fl) and np.array_equal(of.fhigh, fh)
    assert of.get_decs() == fo.get_decs(3)
    A, B, C = fo.weighting_tables(fi, eps=0.0)
    assert np.array_equal(of.A, A) and np.array_equal(of.B, B) and np.array_equal(of.C, C)
    rng = np.random.default_rng(3)
    d0, d1 = rng.standard_normal(24000), rng.standard_normal(24000)
    xr = ref.correlation.generalized_cross_correlation(d0.copy(), d1.copy())
    assert np.array_equal(xr, fo.generalized_cross_correlation(d0, d1))


def test_filter_data_matches_reference(ref):
    """friture_b200/data/filters.npz carries exactly the reference's coefficients."""
    from friture_b200 import filter_data
    P = ref.generated_filters.PARAMS
    bdec, adec, sos = filter_data.decimator()
    assert np.array_equal(bdec, np.array(P["dec"][0])) and np.array_equal(adec, np.array(P["dec"][1]))
    for bpo in (1, 3, 6, 12, 24):
        b, a, s = filter_data.bands(bpo)
        assert np.array_equal(b, np.array(P[str(bpo)][0])) and np.array_equal(a, np.array(P[str(bpo)][1]))


def test_shim_labels_and_attributes_match_reference(ref):
    """Octave_Filters host-side attributes (no GPU needed)."""
    from friture_b200.octavefilters import Octave_Filters
    for bpo in (1, 3, 6, 12, 24):
        r = ref.octavefilters.Octave_Filters(bpo)
        m = Octave_Filters(bpo)
        assert m.f_nominal == r.f_nominal
        for name in ("fi", "flow", "fhigh", "A", "B", "C", "bdec", "adec"):
            assert np.array_equal(getattr(m, name), getattr(r, name)), name
        assert all(np.array_equal(x, y) for x, y in zip(m.boct, r.boct))
        assert all(np.array_equal(x, y) for x, y in zip(m.aoct, r.aoct))
        assert m.get_decs() == r.get_decs() and m.nbands == r.nbands
    from friture_b200 import audioproc
    pr, pm = ref.audioproc.audioproc(), audioproc()
    for n in (1024, 2048):
        pr.set_fftsize(n)
        pm.set_fftsize(n)
        for name in ("window", "freq", "A", "B", "C", "size_sq", "fft_size"):
            assert np.array_equal(getattr(pm, name), getattr(pr, name)), name


def test_live_fft_bank_restatement_and_fir_data(ref):
    """oracle.octave_filter_bank_decimation_fft == the unmodified reference's Octave_Filters.filter
    (live FFT overlap-add path, friture/filter.py:136-247) to 1e-15; the direct FIR form the GPU
    kernel computes equals it to rounding; data/fir.npz holds the reference's taps."""
    from friture_b200 import filter_data
    from oracle import friture_oracle as fo
    assert fo.fft_bank_sizes() == [1536, 1024, 768, 640, 576, 576, 540, 540, 540]
    rng = np.random.default_rng(1)
    for bpo in (1, 3, 24):
        boct_fir, bdec_fir = filter_data.fir_taps(bpo)
        of = ref.octavefilters.Octave_Filters(bpo)
        assert np.array_equal(np.stack(of._boct_fir), boct_fir) and np.array_equal(of._bdec_fir, bdec_fir)
        oo, od = fo.fft_bank_state(bpo)
        hist = [np.zeros(511) for _ in range(9)]
        for blk in range(5):
            x = rng.standard_normal(512 if blk % 2 else 1024) * 0.1
            yr, decr = of.filter(x)
            y, dec, oo, od = fo.octave_filter_bank_decimation_fft(boct_fir, bdec_fir, x, oo, od)
            y2, hist = fo.fir_bank_direct(boct_fir, bdec_fir, x, hist)
            assert dec == decr
            for a, b, c in zip(y, yr, y2):
                assert np.max(np.abs(a - b)) < 1e-14
                assert np.max(np.abs(c - b)) < 1e-12 * max(np.max(np.abs(b)), 1.0)

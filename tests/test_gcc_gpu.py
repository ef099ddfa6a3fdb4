"""GPU parity: GCC-PHAT delay estimation vs the CPU oracle and the reference golden vectors
(friture/signal/correlation.py:24-43; friture/delay_estimator.py:129-152)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ABS_TOL = 1e-4   # identical arg-max, Xcorr within 1e-4 absolute


def load(name):
    with np.load(os.path.join(GOLD, name)) as d:
        return {k: d[k] for k in d.files}


def test_golden_pair_and_smoothing():
    import torch
    from friture_b200.correlation import GccPhat, generalized_cross_correlation
    g = load("gcc_phat.npz")
    x = generalized_cross_correlation(g["d0"], g["d1"])
    assert x.dtype == np.float64 and x.shape == (24000,)
    assert np.max(np.abs(x - g["xcorr"])) < ABS_TOL
    assert int(np.argmax(np.abs(x))) == int(g["argmax"]) == 137
    est = GccPhat(24000)
    d = lambda k: torch.from_numpy(g[k]).cuda()[None, :]
    i1, v1, x1 = est.estimate(d("d0"), d("d1"), want_xcorr=True)
    i2, v2, x2 = est.estimate(d("d0b"), d("d1b"), want_xcorr=True)
    assert int(i1[0]) == 137 and int(i2[0]) == int(g["argmax_b"])
    assert abs(float(v1[0]) - g["xcorr"][137]) < ABS_TOL
    assert abs(float(v2[0]) - g["smoothed_b"][int(g["argmax_b"])]) < ABS_TOL
    assert np.max(np.abs(x2[0].cpu().numpy() - g["xcorr_b"])) < ABS_TOL
    assert np.max(np.abs(est._smoothed[0].cpu().numpy() - g["smoothed_b"])) < ABS_TOL
    assert abs(est.delay_ms(i1)[0] - 1e3 * 137 / 12000.0) < 1e-9


@pytest.mark.parametrize("L", [24000, 12000, 4096, 1000, 6000])
def test_batch_vs_oracle(L):
    import torch
    from friture_b200.correlation import GccPhat
    from oracle import friture_oracle as fo
    rng = np.random.default_rng(L)
    P = 5
    d0 = rng.standard_normal((P, L)).astype(np.float32)
    shifts = [3, L // 7, L // 2 - 1, L - 5, 0]
    d1 = np.stack([np.roll(d0[p], shifts[p]) * (-1 if p == 1 else 1)
                   + 0.1 * rng.standard_normal(L) for p in range(P)]).astype(np.float32)
    est = GccPhat(L)
    idx, val, xc = est.estimate(torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda(),
                                smooth=False, want_xcorr=True)
    xc = xc.cpu().numpy()
    for p in range(P):
        ref = fo.generalized_cross_correlation(d0[p], d1[p])
        assert np.max(np.abs(xc[p] - ref)) < ABS_TOL
        i, v, _ = fo.delay_peak(ref)
        assert int(idx[p]) == i == shifts[p]
        assert abs(float(val[p]) - v) < ABS_TOL
    assert float(val[1]) < 0      # inverted polarity shows in the sign of the extremum


def test_constant_input_and_inputs_untouched():
    import torch
    from friture_b200.correlation import GccPhat
    L = 4096
    d0 = torch.zeros(2, L).cuda()
    d1 = torch.randn(2, L).cuda()
    d1c = d1.clone()
    idx, val, xc = GccPhat(L).estimate(d0, d1, smooth=False, want_xcorr=True)
    assert idx.tolist() == [0, 0] and val.tolist() == [0.0, 0.0]      # delay_estimator.py:164-167
    assert torch.equal(d1, d1c)
    with pytest.raises(ValueError):
        GccPhat(7 * 1024).estimate(torch.zeros(1, 7 * 1024).cuda(), torch.zeros(1, 7 * 1024).cuda())


def test_config4_sample_4096_pairs():
    """BASELINE config #4 geometry on a bounded sample: known integer delays are recovered."""
    import torch
    from friture_b200.correlation import GccPhat
    L, P = 24000, 64
    g = torch.Generator().manual_seed(4)
    d0 = torch.randn(P, L, generator=g)
    k = torch.randint(0, L, (P,), generator=g)
    d1 = torch.stack([torch.roll(d0[p], int(k[p])) for p in range(P)]) + 0.1 * torch.randn(P, L, generator=g)
    idx, val, _ = GccPhat(L).estimate(d0.cuda(), d1.cuda(), smooth=False)
    assert torch.equal(idx.cpu().long(), k)

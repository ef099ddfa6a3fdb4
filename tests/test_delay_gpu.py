"""GPU parity: N-fold decimation and the delay-estimator pipeline vs the CPU oracle
(friture/signal/decimate.py:45-71; friture/delay_estimator.py:87-176)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_decimate_multiple_golden_and_oracle():
    import torch
    from friture_b200 import filter_data
    from friture_b200.delay import Decimator
    from oracle import friture_oracle as fo
    with np.load(os.path.join(GOLD, "gcc_phat.npz")) as g:
        xin, ref = g["dec_in"], g["dec_out"]
    dec = Decimator(1, 2)
    o = torch.cat([dec.process(torch.from_numpy(xin[None, :1024]).cuda()),
                   dec.process(torch.from_numpy(xin[None, 1024:]).cuda())], dim=1)[0].cpu().numpy()
    assert o.shape == ref.shape == (512,)
    assert np.max(np.abs(o - ref)) / np.max(np.abs(ref)) < 5e-5
    # many channels, chunked 512 like the audio backend, one and two stages
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((5, 4096)) * 0.1).astype(np.float32)
    bdec, adec, _ = filter_data.decimator()
    for ns in (1, 2):
        dec = Decimator(5, ns)
        got = torch.cat([dec.process(torch.from_numpy(x[:, p:p + 512]).cuda())
                         for p in range(0, 4096, 512)], dim=1).cpu().numpy()
        for c in range(5):
            zis = [np.zeros(12) for _ in range(ns)]
            want, _ = fo.decimate_multiple(ns, bdec, adec, x[c], zis)
            assert np.max(np.abs(got[c] - want)) / np.max(np.abs(want)) < 5e-5


def test_delay_estimator_pipeline():
    """Two channels, the second delayed by 37 ms: the estimator finds the delay like the
    reference's pipeline restated on the CPU."""
    import torch
    from friture_b200 import filter_data
    from friture_b200.delay import DelayEstimator
    from oracle import friture_oracle as fo
    rng = np.random.default_rng(9)
    fs, n = 48000, 512 * 280
    shift = int(0.037 * fs)
    base = rng.standard_normal(n + shift) * 0.1
    x0 = base[shift:shift + n].astype(np.float32)
    x1 = (base[:n] + 0.01 * rng.standard_normal(n)).astype(np.float32)    # x1 lags x0 by `shift`
    P = 3
    est = DelayEstimator(P, delayrange_s=0.5)
    # CPU restatement of delay_estimator.py:87-152 for one pair
    bdec, adec, _ = filter_data.decimator()
    z0 = [np.zeros(12), np.zeros(12)]
    z1 = [np.zeros(12), np.zeros(12)]
    L, hop = est.length, est.needed
    hist0, hist1 = np.zeros(L), np.zeros(L)
    old_index, offset, old_xc, ref_delay = 0, 0, None, []
    frames = 0
    for p in range(0, n, 512):
        a = torch.from_numpy(np.tile(x0[None, p:p + 512], (P, 1))).cuda()
        b = torch.from_numpy(np.tile(x1[None, p:p + 512], (P, 1))).cuda()
        frames += est.handle_new_data(a, b)
        d0, z0 = fo.decimate_multiple(2, bdec, adec, x0[p:p + 512], z0)
        d1, z1 = fo.decimate_multiple(2, bdec, adec, x1[p:p + 512], z1)
        hist0, hist1 = np.concatenate([hist0, d0]), np.concatenate([hist1, d1])
        offset += len(d0)
        for _ in range(int((offset - old_index) / hop)):
            old_index += hop
            xc = fo.generalized_cross_correlation(hist0[old_index:old_index + L], hist1[old_index:old_index + L])
            i, v, old_xc = fo.delay_peak(xc, old_xc)
            d = 1e3 * i / 12000.0
            ref_delay.append(d - 1e3 if d > 500.0 else d)
    assert frames == len(ref_delay) > 3
    assert np.allclose(est.delay_ms, ref_delay[-1], atol=1e-9)
    assert abs(est.delay_ms[0] - 37.0) < 0.2
    assert np.all(est.correlation > 50)

"""CPU: NumPy float32 emulation of the filterbank kernel's time-parallel biquad evaluation
(friture_b200/csrc/bank.cu: pass 1 zero-state recursion per lane chunk, 5-step inclusive scan
s_l = A^L s_{l-1} + f_l over 32 lanes with precomputed A^(L*2^k), pass 2 from the true incoming
state) against the plain serial recursion.  Documents the algorithm independently of the GPU."""
import numpy as np
import pytest
from scipy.signal import sosfilt

from friture_b200 import filter_data

F = np.float32


def section_chunk_scan(x, sec, carry, L):
    """x float32 [32*L]; sec = (b0,b1,b2,a1,a2) float32; carry = (z1,z2).  Returns (y, new carry)."""
    b0, b1, b2, a1, a2 = [F(v) for v in sec]
    B1, B2 = F(np.float64(b1) - np.float64(a1) * np.float64(b0)), F(np.float64(b2) - np.float64(a2) * np.float64(b0))
    A = np.array([[-np.float64(a1), 1.0], [-np.float64(a2), 0.0]])
    apow = [np.linalg.matrix_power(A, L * 2 ** k).astype(F) for k in range(5)]
    xs = x.reshape(32, L)
    # pass 1: zero-state final state of every lane's chunk
    s = np.zeros((32, 2), dtype=F)
    for l in range(32):
        s1 = s2 = F(0)
        for k in range(L):
            n1 = F(-a1 * s1 + F(B1 * xs[l, k] + s2))
            s2 = F(-a2 * s1 + F(B2 * xs[l, k]))
            s1 = n1
        s[l] = (s1, s2)
    s[0] = s[0] + apow[0] @ np.array(carry, dtype=F)          # fold the carried state into lane 0
    for k in range(5):                                        # Hillis-Steele inclusive scan
        d = 1 << k
        t = s.copy()
        for l in range(d, 32):
            s[l] = t[l] + apow[k] @ t[l - d]
    start = np.vstack([np.array(carry, dtype=F)[None, :], s[:-1]])   # state at each chunk's start
    # pass 2: direct form II transposed from the true incoming state
    y = np.empty_like(xs)
    last = None
    for l in range(32):
        i1, i2 = start[l]
        for k in range(L):
            xx = xs[l, k]
            yy = F(b0 * xx + i1)
            i1 = F(-a1 * yy + F(b1 * xx + i2))
            i2 = F(-a2 * yy + F(b2 * xx))
            y[l, k] = yy
        last = (i1, i2)
    return y.reshape(-1), last


@pytest.mark.parametrize("L", [2, 8, 16])
def test_chunk_scan_equals_serial_recursion(L):
    _, _, sos = filter_data.decimator()
    rng = np.random.default_rng(L)
    x = (rng.standard_normal(3 * 32 * L) * 0.1).astype(F)
    # the whole 6-section decimator over three tiles, state carried between tiles
    carries = [(F(0), F(0))] * 6
    out = []
    for t in range(3):
        cur = x[t * 32 * L:(t + 1) * 32 * L]
        for s in range(6):
            sec = (sos[s, 0], sos[s, 1], sos[s, 2], sos[s, 4], sos[s, 5])
            cur, carries[s] = section_chunk_scan(cur, sec, carries[s], L)
        out.append(cur)
    got = np.concatenate(out).astype(np.float64)
    ref = sosfilt(sos, x.astype(np.float64))
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 2e-5
    # and the float32 serial recursion agrees to float32 rounding
    ser = sosfilt(sos.astype(F), x)
    assert np.max(np.abs(got - ser)) / np.max(np.abs(ref)) < 2e-5

"""Executable model of the lane-pipelined filterbank kernel (friture_b200/csrc/bank_pipe.cu).

TEST INFRASTRUCTURE: a NumPy restatement of the *schedule* the CUDA kernel runs -- which lane
touches which sample at which step, through which shared-memory ring -- with the arithmetic in
float64, so that any difference from the oracle is a scheduling bug, not rounding.  The CUDA
kernel mirrors this file phase by phase (phase A = the lanes' biquad loops, phase B = smoothing
accumulators, block-end reductions, input prefetch, mailbox copy); the tests run it against
oracle/friture_oracle.py on CPU.

Layout of one warp (one channel):
  lanes [0, NSEC)        group 0: section r = lane of stage 0
  lanes [NSEC, 2*NSEC)   group 1: section r of stages >= 1, time-multiplexed: in the CH sample
                         slots of one step, slots [CH-2*len_j, CH-len_j) belong to stage j
                         (len_j = CH >> j, j = 1..LOGCH) and slot CH-1 to the one "ruler" stage
                         JR + ctz(u+1) >= JR = LOGCH+1 that has a sample due
  section r: r < 2*bpo -> band r//2, biquad r%2;  r >= 2*bpo -> decimator biquad r-2*bpo
  skew d(r) = position in its chain: a lane works on the chunk the previous lane of its chain
  finished one step earlier.
"""
import numpy as np

MAX_OCT = 10
DEC_SECTIONS = 6


def stage_start_steps(n_oct, logch):
    """T[j]: step at which the chain heads of stage j work on their first sample/chunk.
    Mirrors pipe_schedule() in bank_pipe.cu."""
    T = [0] * MAX_OCT
    jr = logch + 1
    for j in range(1, MAX_OCT):
        t = T[j - 1] + DEC_SECTIONS
        if j >= jr:
            P = 1 << (j - logch)
            a = P // 2 - 1
            while t % P != a:
                t += 1
        T[j] = t
    return T


def n_steps_for(n_oct, logch, t_total, T):
    ch = 1 << logch
    jr = logch + 1
    n_chunks = t_total // ch
    last = 0
    for j in range(n_oct):
        dmax = 1 if j == n_oct - 1 else DEC_SECTIONS - 1
        if j < jr:
            last = max(last, n_chunks - 1 + T[j] + dmax)
        else:
            m_last = (t_total >> j) - 1
            last = max(last, T[j] + (m_last << (j - logch)) + dmax)
    return last + 1


def ctz(v):
    if v == 0:
        return 32
    n = 0
    while (v & 1) == 0:
        v >>= 1
        n += 1
    return n


class PipeModel:
    def __init__(self, sos_band, sos_dec, alphas, n_oct, logch=5, rx=8, pf=4):
        sos_band = np.asarray(sos_band, dtype=np.float64)
        sos_dec = np.asarray(sos_dec, dtype=np.float64)
        self.bpo = sos_band.shape[0]
        self.n_oct = n_oct
        self.logch = logch
        self.CH = 1 << logch
        self.JR = logch + 1
        self.NSEC = 2 * self.bpo + DEC_SECTIONS
        assert 2 * self.NSEC <= 32
        self.RX, self.PF = rx, pf
        self.alphas = np.ones(MAX_OCT)
        self.alphas[:n_oct] = np.asarray(alphas, dtype=np.float64)[:n_oct]
        self.q = 1.0 - self.alphas
        # normalised sections: numerator (1, c, 1); the chain gain is applied once
        secs = np.concatenate([sos_band.reshape(-1, 6), sos_dec], axis=0)
        assert np.allclose(secs[:, 2], secs[:, 0], rtol=1e-9)
        self.c = secs[:, 1] / secs[:, 0]
        self.a1 = secs[:, 4]
        self.a2 = secs[:, 5]
        self.g_band = np.array([sos_band[i, 0, 0] * sos_band[i, 1, 0] for i in range(self.bpo)])
        self.g_dec = float(np.prod(sos_dec[:, 0]))
        self.T = stage_start_steps(n_oct, logch)
        # canonical state (normalised sections): z[n_oct][NSEC][2], e[n_oct][bpo] (e/alpha form)
        self.z = np.zeros((n_oct, self.NSEC, 2))
        self.e = np.zeros((n_oct, self.bpo))

    # ------------------------------------------------------------------ helpers
    def mux_stage_of_slot(self, p):
        """stage owning slot p of the multiplexed vector (p < CH-1)."""
        CH = self.CH
        j = 1
        while p >= CH - (CH >> j):
            j += 1
        return j

    def process(self, x, block):
        x = np.asarray(x, dtype=np.float64)
        CH, JR, NSEC, bpo, n_oct, logch = self.CH, self.JR, self.NSEC, self.bpo, self.n_oct, self.logch
        RX, PF, T = self.RX, self.PF, self.T
        t_total = x.shape[0]
        assert block & (block - 1) == 0 and block >= 256 and t_total % block == 0
        assert (block >> (n_oct - 1)) >= 1
        n_chunks = t_total // CH
        NB = block // CH
        n_blocks = t_total // block
        n_steps = n_steps_for(n_oct, logch, t_total, T)
        nbands = n_oct * bpo
        energies = np.full((n_blocks, nbands), np.nan)
        NL = 2 * NSEC
        lanes = np.arange(32)
        G = np.where(lanes < NL, lanes // NSEC, 2)
        r = lanes % NSEC
        active_lane = lanes < NL
        isband = r < 2 * bpo
        s = np.where(isband, r % 2, r - 2 * bpo)          # position in the chain = skew d
        d = s
        isdec5 = active_lane & (r == NSEC - 1)
        maxstage = np.where(isband, n_oct - 1, n_oct - 2)
        cc, na1, na2 = self.c[r], -self.a1[r], -self.a2[r]
        # shared memory
        X = np.full((RX, 2 * CH), np.nan)
        L = np.full((32, 2, CH), np.nan)
        S = np.zeros((MAX_OCT, NSEC, 3))                   # z1, z2, e (ruler stages, band s1 roles)
        MB = np.full(MAX_OCT + 1, np.nan)
        S[:n_oct, :, :2] = self.z
        for j in range(JR, n_oct):
            for b in range(bpo):
                S[j, 2 * b + 1, 2] = self.e[j, b]
        # phase-B accumulators, per slot of the two band-output vectors
        acc0 = np.zeros((bpo, CH))
        accm = np.zeros((bpo, CH))
        acc0[:, CH - 1] = self.e[0]
        slot_stage = np.array([self.mux_stage_of_slot(p) for p in range(CH - 1)] + [0])
        for j in range(1, min(JR, n_oct)):
            accm[:, CH - (CH >> j) - 1] = self.e[j]        # last slot of stage j's range
        len_of = lambda j: CH >> j
        # weights: slot p of stage j's range at position pos: q_j^(len_j-1-pos)
        w0 = self.q[0] ** (CH - 1 - np.arange(CH))
        Q0 = self.q[0] ** CH
        wm = np.zeros(CH)
        Qm = np.zeros(CH)
        for p in range(CH - 1):
            j = slot_stage[p]
            if j < n_oct:
                pos = p - (CH - 2 * len_of(j))
                wm[p] = self.q[j] ** (len_of(j) - 1 - pos)
                Qm[p] = self.q[j] ** len_of(j)
        # prologue: prefetch chunks 0..PF-1
        for cpre in range(min(PF, n_chunks)):
            X[cpre % RX, :CH] = x[cpre * CH:(cpre + 1) * CH]

        def band_out(stage, b, val):
            k = (n_oct - 1 - stage) * bpo + b
            return k, self.alphas[stage] * val

        for k in range(n_steps):
            # ---------------------------------------------------------------- phase A
            Lnew = {}
            Xw = {}
            MBw = {}
            for lane in range(NL):
                g, rr, dd = G[lane], r[lane], d[lane]
                u = k - dd
                if s[lane] == 0:
                    inp = X[k % RX, g * CH:(g + 1) * CH]
                else:
                    inp = L[lane - 1, (k - 1) & 1]
                out = np.full(CH, np.nan)
                # segments: (slot_start, length, stage, chunk index, valid)
                segs = []
                for gg in range(logch):            # len CH/2 ... 1
                    ln = CH >> (gg + 1)
                    st = CH - 2 * ln
                    stage = g * (gg + 1)
                    cidx = u - (T[stage] if g else 0)
                    valid = (0 <= cidx < n_chunks) and stage <= maxstage[lane]
                    segs.append((st, ln, stage, cidx, valid))
                # ruler slot
                if g == 0:
                    segs.append((CH - 1, 1, 0, u, (0 <= u < n_chunks) and 0 <= maxstage[lane]))
                    jr, m = 0, u
                else:
                    jr = JR + ctz(u + 1) if u >= 0 else 99
                    valid = False
                    m = -1
                    if jr <= maxstage[lane] and u >= T[jr]:
                        m = (u - T[jr]) >> (jr - logch)
                        valid = m < (t_total >> jr)
                    segs.append((CH - 1, 1, min(jr, MAX_OCT - 1), m, valid))
                for (st, ln, stage, cidx, valid) in segs:
                    z1, z2, e = S[stage, rr]
                    for i in range(st, st + ln):
                        xv = inp[i]
                        y = xv + z1
                        z1 = na1[lane] * y + (cc[lane] * xv + z2)
                        z2 = na2[lane] * y + xv
                        out[i] = y
                    if st == CH - 1 and g == 1 and isband[lane] and s[lane] == 1:
                        # ruler stage smoothing in the lane (e/alpha form)
                        b = rr // 2
                        yy = out[CH - 1] * self.g_band[b]
                        e = e * self.q[stage] + yy * yy
                        if valid and ((cidx + 1) & ((block >> stage) - 1)) == 0:
                            kb, val = band_out(stage, b, e)
                            energies[(cidx + 1) // (block >> stage) - 1, kb] = val
                    if valid:
                        S[stage, rr] = (z1, z2, e)
                if isdec5[lane]:
                    base = CH + g * (CH // 2)
                    for i in range(0, CH - 4, 2):
                        Xw[((k + 1) % RX, base + i // 2)] = out[i] * self.g_dec
                    Xw[((k + 1) % RX, base + (CH - 4) // 2)] = out[CH - 4] * self.g_dec
                    if g == 0:
                        Xw[((k + 1) % RX, base + (CH - 2) // 2)] = out[CH - 2] * self.g_dec
                    else:
                        st, ln, stage, cidx, valid = segs[logch - 1]     # the len-1 stage JR-1
                        if valid and (cidx % 2 == 0):
                            MBw[JR] = out[CH - 2] * self.g_dec
                        st, ln, stage, cidx, valid = segs[logch]
                        if valid and (cidx % 2 == 0):
                            MBw[stage + 1] = out[CH - 1] * self.g_dec
                else:
                    Lnew[lane] = out
            for lane, out in Lnew.items():
                L[lane, k & 1] = out
            for (slot, pos), v in Xw.items():
                X[slot, pos] = v
            for j, v in MBw.items():
                MB[j] = v
            # ---------------------------------------------------------------- phase B
            c0 = k - 1
            valid0 = 0 <= c0 < n_chunks
            for b in range(bpo):
                y0 = L[2 * b + 1, k & 1] * self.g_band[b]
                ym = L[NSEC + 2 * b + 1, k & 1] * self.g_band[b]
                if valid0:
                    acc0[b] = acc0[b] * Q0 + y0 * y0
                    if ((c0 + 1) & (NB - 1)) == 0:
                        tot = float(np.dot(w0, acc0[b]))
                        kb, val = band_out(0, b, tot)
                        energies[(c0 + 1) // NB - 1, kb] = val
                        acc0[b] = 0.0
                        acc0[b, CH - 1] = tot
                for j in range(1, min(JR, n_oct)):
                    cm = k - 1 - T[j]
                    if not (0 <= cm < n_chunks):
                        continue
                    lo, hi = CH - 2 * len_of(j), CH - len_of(j)
                    accm[b, lo:hi] = accm[b, lo:hi] * Qm[lo:hi] + ym[lo:hi] ** 2
                    if ((cm + 1) & (NB - 1)) == 0:
                        tot = float(np.dot(wm[lo:hi], accm[b, lo:hi]))
                        kb, val = band_out(j, b, tot)
                        energies[(cm + 1) // NB - 1, kb] = val
                        accm[b, lo:hi] = 0.0
                        accm[b, hi - 1] = tot
            # prefetch chunk k + PF, mailbox -> X for the next step
            cn = k + PF
            if cn < n_chunks:
                X[cn % RX, :CH] = x[cn * CH:(cn + 1) * CH]
            jn = JR + ctz(k + 2)
            if jn <= n_oct - 1:
                X[(k + 1) % RX, 2 * CH - 1] = MB[jn]
        # epilogue: canonical state
        self.z = S[:n_oct, :, :2].copy()
        self.e[0] = acc0[:, CH - 1]
        for j in range(1, min(JR, n_oct)):
            self.e[j] = accm[:, CH - len_of(j) - 1]
        for j in range(JR, n_oct):
            for b in range(bpo):
                self.e[j, b] = S[j, 2 * b + 1, 2]
        assert not np.isnan(energies).any(), "some band energies were never emitted"
        return energies

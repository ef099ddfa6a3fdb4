// Lane-pipelined multi-rate IIR filterbank + exponential RMS for sm_100a (bpo = 1 or 3).
//
// Same numerics contract as bank.cu (the reference's IIR bank octave_filter_bank_decimation,
// friture/filter.py:86-118, recursion friture/signal/lfilter.py:131-139, [::2] of
// friture/signal/decimate.py:39-41, fused with exp_smoothed_value(y**2),
// friture/octavespectrum.py:104 / friture/signal/exp_smoothing.py:11-56), different schedule:
//
// ONE WARP PER CHANNEL (or per channel PAIR, two channels packed in float2 -> FFMA2), and inside
// the warp ONE LANE PER BIQUAD SECTION.  A recursion is serial in time, so instead of splitting
// time (bank.cu: two passes + a scan per section) every lane runs ITS section serially over a
// chunk of CH samples per step and hands the chunk to the next section of its chain through a
// shared-memory ring; the chain is a software pipeline, lane r works on the chunk lane r-1
// finished one step earlier.  Each sample of each section is computed exactly once, with the
// 4-operation normalised biquad (bank_internal.cuh).
//
//   lanes [0, NSEC)       sections of stage 0 (rate fs):  bands (2 sections each), 6 decimator
//   lanes [NSEC, 2 NSEC)  the same sections for ALL lower-rate stages, time-multiplexed: of the
//                         CH sample slots of a step, slots [CH-2 len_j, CH-len_j) belong to stage
//                         j (len_j = CH >> j) and the last slot to the one stage >= JR = log2(CH)+1
//                         whose turn it is (stage JR + ctz(u+1): a binary-ruler schedule, every
//                         stage gets exactly its 2^-j share) -- 31/32 of these lanes' slots are used.
//   The last decimator lane keeps the even samples (x gain) and writes them where the next
//   stage's chain heads will read them one step later (slot i -> CH/2 + i/2 of the multiplexed
//   input vector; single samples of the ruler stages wait in a per-stage mailbox).
//   Smoothing of y^2 is NOT done in the section loops: after each step all 32 lanes update one
//   accumulator per sample slot (acc <- q^len acc + y^2, 3 instructions per band per step instead
//   of 2 per sample); at a block end the accumulators of a stage are combined with a weighted
//   (segmented) warp reduction into the band energy and collapsed back to one value, so the state
//   carried between launches is the plain smoothed energy.  Decays are applied in complement form
//   (acc - (1-q^n) acc) so that the float32 rounding of q does not bias long time constants.
//
// tests/bank_pipeline_model.py is an executable NumPy model of exactly this schedule, checked
// against the oracle on CPU; this file mirrors it phase by phase.
#include <cmath>
#include <cstdlib>

#include "bank_internal.cuh"

namespace {

constexpr int DEC_SECTIONS = 6;

template <int PACK> struct VT;
template <> struct VT<1> { using t = float; };
template <> struct VT<2> { using t = float2; };

__device__ __forceinline__ float v_add(float a, float b) { return a + b; }
__device__ __forceinline__ float2 v_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float v_fma(float c, float x, float z) { return fmaf(c, x, z); }
__device__ __forceinline__ float2 v_fma(float c, float2 x, float2 z) {
    return __ffma2_rn(make_float2(c, c), x, z);
}
__device__ __forceinline__ float v_mul(float c, float x) { return c * x; }
__device__ __forceinline__ float2 v_mul(float c, float2 x) { return __fmul2_rn(make_float2(c, c), x); }
__device__ __forceinline__ float v_sq(float a) { return a * a; }
__device__ __forceinline__ float2 v_sq(float2 a) { return __fmul2_rn(a, a); }
__device__ __forceinline__ void v_set(float &d, float a, float) { d = a; }
__device__ __forceinline__ void v_set(float2 &d, float a, float b) { d = make_float2(a, b); }
__device__ __forceinline__ float v_x(float a) { return a; }
__device__ __forceinline__ float v_x(float2 a) { return a.x; }
__device__ __forceinline__ float v_y(float a) { return a; }
__device__ __forceinline__ float v_y(float2 a) { return a.y; }
__device__ __forceinline__ float v_shfl_xor(float a, int d) { return __shfl_xor_sync(0xffffffffu, a, d); }
__device__ __forceinline__ float2 v_shfl_xor(float2 a, int d) {
    return make_float2(__shfl_xor_sync(0xffffffffu, a.x, d), __shfl_xor_sync(0xffffffffu, a.y, d));
}

__device__ __forceinline__ void ld4(const float *p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const float2 *p, float2 (&v)[4]) {
    const float4 a = reinterpret_cast<const float4 *>(p)[0];
    const float4 b = reinterpret_cast<const float4 *>(p)[1];
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
    v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void st4(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(float2 *p, const float2 (&v)[4]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}
__device__ __forceinline__ void st2(float *p, float a, float b) {
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
}
__device__ __forceinline__ void st2(float2 *p, float2 a, float2 b) {
    *reinterpret_cast<float4 *>(p) = make_float4(a.x, a.y, b.x, b.y);
}
// (z1, z2) of one section: 2 T's, 8*PACK bytes, aligned
__device__ __forceinline__ void ldz(const float *p, float &z1, float &z2) {
    const float2 t = *reinterpret_cast<const float2 *>(p);
    z1 = t.x; z2 = t.y;
}
__device__ __forceinline__ void ldz(const float2 *p, float2 &z1, float2 &z2) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    z1 = make_float2(t.x, t.y); z2 = make_float2(t.z, t.w);
}
__device__ __forceinline__ void stz(float *p, float z1, float z2) {
    *reinterpret_cast<float2 *>(p) = make_float2(z1, z2);
}
__device__ __forceinline__ void stz(float2 *p, float2 z1, float2 z2) {
    *reinterpret_cast<float4 *>(p) = make_float4(z1.x, z1.y, z2.x, z2.y);
}

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float lg2_fast(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// friture/octavespectrum.py:119-121: 10*log10(sp + 1e-30) + w
__device__ __forceinline__ float energy_out(float e, int kband, const BankArgs &a) {
    if (!a.db) return e;
    float v = 3.01029995663981195f * lg2_fast(e + 1e-30f);
    if (a.weight) v += __ldg(a.weight + kband);
    return v;
}

template <class T>
__device__ __forceinline__ void emit(const BankArgs &a, float alpha_j, T val, int ch0, bool has2,
                                     int blk, int kband, int nbands) {
    if (!a.energies) return;
    float *o = a.energies + ((size_t)ch0 * a.n_blocks + blk) * nbands + kband;
    o[0] = energy_out(alpha_j * v_x(val), kband, a);
    if (sizeof(T) == 8 && has2) o[(size_t)a.n_blocks * nbands] = energy_out(alpha_j * v_y(val), kband, a);
}

// normalised biquad step (bank_internal.cuh): y = x + z1; z1 = c x - a1 y + z2; z2 = x - a2 y
template <class T>
__device__ __forceinline__ T biquad(T x, T &z1, T &z2, float cc, float na1, float na2) {
    const T y = v_add(x, z1);
    const T t = v_fma(cc, x, z2);
    z1 = v_fma(na1, y, t);
    z2 = v_fma(na2, y, x);
    return y;
}

template <int LOGCH, int PACK, int BPO>
__global__ void __launch_bounds__(32)
bank_pipe_kernel(const __grid_constant__ PipeParams P, const BankArgs a) {
    using T = typename VT<PACK>::t;
    constexpr int CH = 1 << LOGCH, JR = LOGCH + 1, NSEC = 2 * BPO + DEC_SECTIONS, NL = 2 * NSEC;
    constexpr int SPL = CH / 32;
    constexpr int RX = PIPE_RX, PF = PIPE_PF;
    constexpr int LPAD = 16 / (int)sizeof(T);
    constexpr int LSTR = 2 * CH + LPAD;       // per-lane output ring: 2 chunks + 16 B of bank skew
    static_assert(NL <= 32, "two lane groups must fit in a warp");
    static_assert(LOGCH == 5 || LOGCH == 6, "steps of 32 or 64 samples");
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ float4 smem4[];
    T *sX = reinterpret_cast<T *>(smem4);             // [RX][2 CH]: stage-0 chunk | multiplexed vector
    T *sL = sX + RX * 2 * CH;                         // [NL][LSTR]
    T *sS = sL + NL * LSTR;                           // [MAX_OCT][NSEC][4]: z1 z2 e -
    T *sMB = sS + BANK_MAX_OCT * NSEC * 4;            // [MAX_OCT + 1] mailboxes of the ruler stages
    int *sT = reinterpret_cast<int *>(sMB + 12);      // [MAX_OCT + 1]
    float *sAl = reinterpret_cast<float *>(sT + 12);  // [MAX_OCT + 1]

    const int lane = threadIdx.x;
    const int ch0 = blockIdx.x * PACK;
    const bool has2 = (PACK == 2) && (ch0 + 1 < a.n_channels);
    const int ch1 = has2 ? ch0 + 1 : ch0;
    const int n_oct = P.n_oct;
    const int nbands = n_oct * BPO;
    const int n_chunks = (int)(a.t_total >> LOGCH);
    const int lognb = 31 - __clz(a.block) - LOGCH;          // block = CH << lognb
    const int nbmask = (1 << lognb) - 1;
    const int logblock = lognb + LOGCH;

    // ---- lane roles
    const bool worker = lane < NL;
    const int G = worker ? lane / NSEC : 0;
    const int r = worker ? lane - G * NSEC : 0;
    const bool isband = r < 2 * BPO;
    const int s = isband ? (r & 1) : r - 2 * BPO;            // position in the chain = skew
    const bool isdec5 = (r == NSEC - 1);
    const int maxstage = isband ? n_oct - 1 : n_oct - 2;     // the last stage's decimator is unused (filter.py:113)
    const bool isemar = (G == 1) && isband && (s == 1);      // smooths the ruler stages in the lane
    const float cc = P.c[r], na1 = P.na1[r], na2 = P.na2[r];
    const float gb_lane = isband ? P.gband[r >> 1] : 0.f;
    const float gdec = P.gdec;

    // ---- prologue: tables, state
    if (lane <= BANK_MAX_OCT) {
        sT[lane] = P.T[lane];
        sAl[lane] = P.alpha[lane];
    }
    const float *gz0 = a.zstate + (size_t)ch0 * n_oct * NSEC * 2;
    const float *gz1 = a.zstate + (size_t)ch1 * n_oct * NSEC * 2;
    float *ge0 = a.ema + (size_t)ch0 * nbands;
    float *ge1 = a.ema + (size_t)ch1 * nbands;
    for (int i = lane; i < BANK_MAX_OCT * NSEC; i += 32) {
        const int j = i / NSEC, rr = i - j * NSEC;
        T z1, z2, e;
        v_set(z1, 0.f, 0.f); v_set(z2, 0.f, 0.f); v_set(e, 0.f, 0.f);
        if (j < n_oct) {
            v_set(z1, gz0[2 * i], gz1[2 * i]);
            v_set(z2, gz0[2 * i + 1], gz1[2 * i + 1]);
            if (j >= JR && rr < 2 * BPO && (rr & 1)) v_set(e, ge0[j * BPO + (rr >> 1)], ge1[j * BPO + (rr >> 1)]);
        }
        sS[i * 4 + 0] = z1; sS[i * 4 + 1] = z2; sS[i * 4 + 2] = e; sS[i * 4 + 3] = e;
    }
    // smoothing accumulators: one per sample slot (slot = lane + 32 s) of the two band-output vectors
    T acc0[BPO][SPL], accm[BPO][SPL];
    int Tm[SPL], mst[SPL], gsz[SPL];
    bool mok[SPL], lastslot[SPL], firstslot[SPL];
    float om0[SPL], aqm[SPL], omm[SPL];
#pragma unroll
    for (int q = 0; q < SPL; q++) {
        const int p = lane + 32 * q;
        const int j = 1 + __clz(~((unsigned)p << (32 - LOGCH)));     // stage owning slot p
        mst[q] = j;
        mok[q] = (j < JR) && (j <= n_oct - 1);
        gsz[q] = CH >> j;
        Tm[q] = P.T[j < BANK_MAX_OCT ? j : BANK_MAX_OCT];
        firstslot[q] = (p == CH - 2 * (CH >> j));
        lastslot[q] = (p == CH - (CH >> j) - 1);
        om0[q] = P.om0[q][lane];
        aqm[q] = P.aqm[q][lane];
        omm[q] = P.omm[q][lane];
#pragma unroll
        for (int b = 0; b < BPO; b++) {
            v_set(acc0[b][q], 0.f, 0.f);
            v_set(accm[b][q], 0.f, 0.f);
            if (p == CH - 1) v_set(acc0[b][q], ge0[b], ge1[b]);
            if (mok[q] && lastslot[q]) v_set(accm[b][q], ge0[j * BPO + b], ge1[j * BPO + b]);
        }
    }
    const float *x0 = a.x + (size_t)ch0 * a.x_stride;
    const float *x1 = a.x + (size_t)ch1 * a.x_stride;
    auto prefetch = [&](int cn) {
        if (cn < n_chunks) {
            T *dst = sX + (cn & (RX - 1)) * 2 * CH;
#pragma unroll
            for (int q = 0; q < SPL; q++) {
                const int p = lane + 32 * q;
                float *d = reinterpret_cast<float *>(dst + p);
                cp_async4(d, x0 + (size_t)cn * CH + p);
                if (PACK == 2) cp_async4(d + 1, x1 + (size_t)cn * CH + p);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int cn = 0; cn < PF; cn++) prefetch(cn);
    cp_async_wait<PF - 1>();
    __syncwarp();

    // stage-0 lanes carry their section state in registers from step to step
    T zc1 = sS[r * 4 + 0], zc2 = sS[r * 4 + 1];
    const int sstep = G ? NSEC * 4 : 0;                     // T elements per stage row of sS
    const int tstep = G ? DEC_SECTIONS : 0;                 // T[j] = 6 j for the chunked stages
    const int n_steps = a.n_steps;

    for (int k = 0; k < n_steps; k++) {
        // ============================================================ phase A: the section loops
        if (worker) {
            const int u = k - s;
            const T *inp = (s == 0) ? sX + (k & (RX - 1)) * 2 * CH + G * CH
                                    : sL + (lane - 1) * LSTR + ((k - 1) & 1) * CH;
            T *outp = sL + lane * LSTR + (k & 1) * CH;
            T *xn = sX + ((k + 1) & (RX - 1)) * 2 * CH + CH + G * (CH / 2);
            T *sp = sS + r * 4;
            int cidx = u;
            T z1 = zc1, z2 = zc2;
            const bool pv0 = (unsigned)u < (unsigned)n_chunks && 0 <= maxstage;
            // ---- segments of 4 or more samples: stage G*(g+1), slots [CH - 2 len, CH - len)
#pragma unroll
            for (int g = 0; g <= LOGCH - 3; g++) {
                const int len = CH >> (g + 1), off = CH - 2 * len;
                sp += sstep;
                cidx -= tstep;
                const bool pv = (unsigned)cidx < (unsigned)n_chunks && (G ? g + 1 : 0) <= maxstage;
                if (G) ldz(sp, z1, z2);
#pragma unroll
                for (int q = 0; q < len / 4; q++) {
                    T v[4], y[4];
                    ld4(inp + off + 4 * q, v);
#pragma unroll
                    for (int i = 0; i < 4; i++) y[i] = biquad(v[i], z1, z2, cc, na1, na2);
                    if (!isdec5) st4(outp + off + 4 * q, y);
                    else st2(xn + (off + 4 * q) / 2, v_mul(gdec, y[0]), v_mul(gdec, y[2]));
                }
                if (G && pv) stz(sp, z1, z2);
            }
            // ---- last group of 4 slots: a 2-sample stage, a 1-sample stage, the ruler slot
            {
                T v[4], y[4];
                ld4(inp + CH - 4, v);
                sp += sstep;
                cidx -= tstep;
                const bool pvA = (unsigned)cidx < (unsigned)n_chunks && (G ? LOGCH - 1 : 0) <= maxstage;
                if (G) ldz(sp, z1, z2);
                y[0] = biquad(v[0], z1, z2, cc, na1, na2);
                y[1] = biquad(v[1], z1, z2, cc, na1, na2);
                if (G && pvA) stz(sp, z1, z2);
                sp += sstep;
                cidx -= tstep;
                const bool pvB = (unsigned)cidx < (unsigned)n_chunks && (G ? LOGCH : 0) <= maxstage;
                const int cB = cidx;
                if (G) ldz(sp, z1, z2);
                y[2] = biquad(v[2], z1, z2, cc, na1, na2);
                if (G && pvB) stz(sp, z1, z2);
                // ruler slot: the stage >= JR whose sample is due (group 1); stage 0 for group 0
                int jrc = 0, m = u;
                bool pvR = pv0;
                T e;
                v_set(e, 0.f, 0.f);
                T *spr = sS + r * 4;
                if (G) {
                    const int jr = JR - 1 + __ffs(u + 1);
                    jrc = jr < BANK_MAX_OCT - 1 ? jr : BANK_MAX_OCT - 1;
                    const int Tj = sT[jrc];
                    m = (u - Tj) >> (jrc - LOGCH);
                    pvR = jr <= maxstage && u >= Tj && m < (int)(a.t_total >> jrc);
                    spr = sS + (jrc * NSEC + r) * 4;
                    ldz(spr, z1, z2);
                    e = spr[2];
                }
                y[3] = biquad(v[3], z1, z2, cc, na1, na2);
                if (G && pvR) stz(spr, z1, z2);
                if (isemar) {
                    // exp_smoothed_value of the low-rate stages: e <- (1-alpha) e + y^2 (e/alpha form)
                    const float alj = sAl[jrc];
                    const T yy = v_mul(gb_lane, y[3]);
                    e = v_add(v_fma(-alj, e, e), v_sq(yy));
                    if (pvR) {
                        spr[2] = e;
                        const int bl = logblock - jrc;           // block >> jrc = 2^bl samples
                        if (((m + 1) & ((1 << bl) - 1)) == 0)
                            emit<T>(a, alj, e, ch0, has2, ((m + 1) >> bl) - 1,
                                    (n_oct - 1 - jrc) * BPO + (r >> 1), nbands);
                    }
                }
                if (!isdec5) {
                    st4(outp + CH - 4, y);
                } else {
                    xn[(CH - 4) / 2] = v_mul(gdec, y[0]);
                    if (!G) {
                        xn[(CH - 2) / 2] = v_mul(gdec, y[2]);
                    } else {
                        if (pvB && !(cB & 1)) sMB[JR] = v_mul(gdec, y[2]);
                        if (pvR && !(m & 1)) sMB[jrc + 1] = v_mul(gdec, y[3]);
                    }
                }
                if (!G && pv0) { zc1 = z1; zc2 = z2; }
            }
        }
        __syncwarp();
        // ============================================================ phase B: smoothing, prefetch
        {
            const int c0i = k - 1;
            const bool valid0 = (unsigned)c0i < (unsigned)n_chunks;
            const bool end0 = valid0 && (((c0i + 1) & nbmask) == 0);
            bool vm[SPL], em[SPL];
            int cmv[SPL];
            bool anyend = false;
#pragma unroll
            for (int q = 0; q < SPL; q++) {
                cmv[q] = k - 1 - Tm[q];
                vm[q] = mok[q] && (unsigned)cmv[q] < (unsigned)n_chunks;
                em[q] = vm[q] && (((cmv[q] + 1) & nbmask) == 0);
                anyend = anyend || em[q];
            }
            const int buf = (k & 1) * CH;
#pragma unroll
            for (int b = 0; b < BPO; b++) {
                const T *y0p = sL + (2 * b + 1) * LSTR + buf;
                const T *ymp = sL + (NSEC + 2 * b + 1) * LSTR + buf;
                const float gb = P.gband[b];
#pragma unroll
                for (int q = 0; q < SPL; q++) {
                    const int p = lane + 32 * q;
                    const T yy = v_mul(gb, y0p[p]);
                    const T ym = v_mul(gb, ymp[p]);
                    if (valid0) acc0[b][q] = v_add(v_fma(-P.aq0, acc0[b][q], acc0[b][q]), v_sq(yy));
                    if (vm[q]) accm[b][q] = v_add(v_fma(-aqm[q], accm[b][q], accm[b][q]), v_sq(ym));
                }
            }
            if (end0) {     // stage 0 finished a block: weighted sum of its CH accumulators
                const int blk = ((c0i + 1) >> lognb) - 1;
#pragma unroll
                for (int b = 0; b < BPO; b++) {
                    T val = v_fma(-om0[0], acc0[b][0], acc0[b][0]);
#pragma unroll
                    for (int q = 1; q < SPL; q++) val = v_add(val, v_fma(-om0[q], acc0[b][q], acc0[b][q]));
#pragma unroll
                    for (int dlt = 16; dlt >= 1; dlt >>= 1) val = v_add(val, v_shfl_xor(val, dlt));
#pragma unroll
                    for (int q = 0; q < SPL; q++) {
                        v_set(acc0[b][q], 0.f, 0.f);
                        if (lane + 32 * q == CH - 1) acc0[b][q] = val;
                    }
                    if (lane == 0) emit<T>(a, P.alpha[0], val, ch0, has2, blk, (n_oct - 1) * BPO + b, nbands);
                }
            }
            if (__any_sync(FULL, anyend)) {   // some lower-rate stage finished a block
#pragma unroll
                for (int b = 0; b < BPO; b++) {
#pragma unroll
                    for (int q = 0; q < SPL; q++) {
                        T val;
                        v_set(val, 0.f, 0.f);
                        if (em[q]) val = v_fma(-omm[q], accm[b][q], accm[b][q]);
#pragma unroll
                        for (int dlt = 1; dlt <= 16; dlt <<= 1) {
                            const T t = v_shfl_xor(val, dlt);
                            if (gsz[q] > dlt) val = v_add(val, t);
                        }
                        if (em[q]) {
                            v_set(accm[b][q], 0.f, 0.f);
                            if (lastslot[q]) accm[b][q] = val;
                            if (firstslot[q])
                                emit<T>(a, sAl[mst[q]], val, ch0, has2, ((cmv[q] + 1) >> lognb) - 1,
                                        (n_oct - 1 - mst[q]) * BPO + b, nbands);
                        }
                    }
                }
            }
            prefetch(k + PF);
            cp_async_wait<PF - 1>();           // chunk k+1 has landed
            if (lane == 0) {                   // the ruler stage of step k+1 reads its sample from the ring
                const int jn = JR - 1 + __ffs(k + 2);
                if (jn <= n_oct - 1) sX[((k + 1) & (RX - 1)) * 2 * CH + 2 * CH - 1] = sMB[jn];
            }
        }
        __syncwarp();
    }

    // ---- epilogue: the pipeline is drained, every stage ended on a block boundary
    if (worker && !G) {
        sS[r * 4 + 0] = zc1;
        sS[r * 4 + 1] = zc2;
    }
    __syncwarp();
    float *wz0 = a.zstate + (size_t)ch0 * n_oct * NSEC * 2;
    float *wz1 = a.zstate + (size_t)ch1 * n_oct * NSEC * 2;
    for (int i = lane; i < n_oct * NSEC; i += 32) {
        const T z1 = sS[i * 4 + 0], z2 = sS[i * 4 + 1];
        wz0[2 * i] = v_x(z1);
        wz0[2 * i + 1] = v_x(z2);
        if (has2) {
            wz1[2 * i] = v_y(z1);
            wz1[2 * i + 1] = v_y(z2);
        }
    }
#pragma unroll
    for (int q = 0; q < SPL; q++) {
        const int p = lane + 32 * q;
#pragma unroll
        for (int b = 0; b < BPO; b++) {
            if (p == CH - 1) {
                ge0[b] = v_x(acc0[b][q]);
                if (has2) ge1[b] = v_y(acc0[b][q]);
            }
            if (mok[q] && lastslot[q]) {
                ge0[mst[q] * BPO + b] = v_x(accm[b][q]);
                if (has2) ge1[mst[q] * BPO + b] = v_y(accm[b][q]);
            }
        }
    }
    for (int i = lane; i < (n_oct - JR) * BPO; i += 32) {     // ruler stages
        const int j = JR + i / BPO, b = i % BPO;
        const T e = sS[(j * NSEC + 2 * b + 1) * 4 + 2];
        ge0[j * BPO + b] = v_x(e);
        if (has2) ge1[j * BPO + b] = v_y(e);
    }
}

template <int LOGCH, int PACK, int BPO>
cudaError_t launch_pipe(const PipeParams &P, const BankArgs &a, cudaStream_t st) {
    constexpr int CH = 1 << LOGCH, NSEC = 2 * BPO + DEC_SECTIONS, NL = 2 * NSEC;
    constexpr size_t TS = 4 * PACK;
    constexpr size_t LSTR = 2 * CH + 16 / TS;
    const size_t smem = TS * (PIPE_RX * 2 * CH + NL * LSTR + BANK_MAX_OCT * NSEC * 4 + 12) + 2 * 12 * 4;
    auto kern = bank_pipe_kernel<LOGCH, PACK, BPO>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((a.n_channels + PACK - 1) / PACK);
    kern<<<blocks, 32, smem, st>>>(P, a);
    return cudaGetLastError();
}

template <int LOGCH, int PACK>
cudaError_t launch_pipe_bpo(const PipeParams &P, const BankArgs &a, cudaStream_t st) {
    if (P.bpo == 3) return launch_pipe<LOGCH, PACK, 3>(P, a, st);
    return launch_pipe<LOGCH, PACK, 1>(P, a, st);
}

}   // namespace

// Step at which the chain heads of stage j start (T[j]) and the number of steps that drains the
// pipeline for t_total samples per channel.  Stages j <= logch move a chunk of 2^logch >> j samples
// per step, six steps (the decimator chain) behind the previous stage; stage j > logch has one
// sample every P_j = 2^(j-logch) steps, at steps u = T_j (mod P_j) with T_j = P_j/2 - 1 (mod P_j),
// which makes the stages' turns disjoint (the ruler sequence JR + ctz(u+1)).
void frt_pipe_schedule(int n_oct, int logch, long long t_total, int *T, int *n_steps) {
    T[0] = 0;
    for (int j = 1; j <= BANK_MAX_OCT; j++) {
        int t = T[j - 1] + DEC_SECTIONS;
        if (j > logch) {
            const int P = 1 << (j - logch), arem = P / 2 - 1;
            while (t % P != arem) t++;
        }
        T[j] = t;
    }
    if (!n_steps) return;
    const long long n_chunks = t_total >> logch;
    long long last = 0;
    for (int j = 0; j < n_oct; j++) {
        const int dmax = (j == n_oct - 1) ? 1 : DEC_SECTIONS - 1;
        long long l;
        if (j <= logch) l = n_chunks - 1 + T[j] + dmax;
        else l = T[j] + (((t_total >> j) - 1) << (j - logch)) + dmax;
        if (l > last) last = l;
    }
    *n_steps = (int)(last + 1);
}

void frt_pipe_prepare(BankPlan *pl) {
    const BankParams &B = pl->params;
    pl->pipe_ok = (B.bpo == 1 || B.bpo == 3);
    if (!pl->pipe_ok) return;
    for (int v = 0; v < 2; v++) {
        const int logch = 5 + v, CH = 1 << logch;
        PipeParams &P = pl->pipe[v];
        memset(&P, 0, sizeof(P));
        P.n_oct = B.n_oct;
        P.bpo = B.bpo;
        for (int r = 0; r < B.nsec; r++) {
            P.c[r] = B.coef[r][1];
            P.na1[r] = -B.coef[r][3];
            P.na2[r] = -B.coef[r][4];
        }
        for (int b = 0; b < B.bpo; b++) P.gband[b] = B.gband[b];
        P.gdec = B.gdec;
        for (int j = 0; j <= BANK_MAX_OCT; j++) P.alpha[j] = j < B.n_oct ? (float)pl->alphas[j] : 1.f;
        frt_pipe_schedule(B.n_oct, logch, 0, P.T, nullptr);
        const double q0 = 1.0 - pl->alphas[0];
        P.aq0 = (float)(1.0 - pow(q0, CH));
        for (int p = 0; p < CH; p++) {
            P.om0[p >> 5][p & 31] = (float)(1.0 - pow(q0, CH - 1 - p));
            int j = 1;
            while (j <= logch && p >= CH - (CH >> j)) j++;
            if (j <= logch && j < B.n_oct) {
                const int len = CH >> j, pos = p - (CH - 2 * len);
                const double qj = 1.0 - pl->alphas[j];
                P.aqm[p >> 5][p & 31] = (float)(1.0 - pow(qj, len));
                P.omm[p >> 5][p & 31] = (float)(1.0 - pow(qj, len - 1 - pos));
            }
        }
    }
}

cudaError_t frt_pipe_launch(const BankPlan *pl, BankArgs a, int logch, int pack, cudaStream_t st) {
    const PipeParams &P = pl->pipe[logch - 5];
    int T[BANK_MAX_OCT + 1];
    frt_pipe_schedule(P.n_oct, logch, a.t_total, T, &a.n_steps);
    if (logch == 5) return pack == 2 ? launch_pipe_bpo<5, 2>(P, a, st) : launch_pipe_bpo<5, 1>(P, a, st);
    return pack == 2 ? launch_pipe_bpo<6, 2>(P, a, st) : launch_pipe_bpo<6, 1>(P, a, st);
}
